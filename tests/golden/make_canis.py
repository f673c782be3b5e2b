#!/usr/bin/env python
"""Builds tests/golden/canis_c1.json from the reference's tutorial data (BASELINE.json configs[0]: the Canis
multi-locus analysis, tutorial/CanisPhylogeny/Canis-Phylogeny.xml: 16 loci x 16 individuals / 8 species, HKY, constant
populations, strict clock).  Only the compressed site patterns travel (a few KB); run here, where /root/reference exists:
    python tests/golden/make_canis.py
"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from starbeast2_b200.io import read_fasta  # noqa: E402
from starbeast2_b200.plugins import Alignment  # noqa: E402

DATA = "/root/reference/tutorial/data"


def main():
    loci = {}
    for path in sorted(glob.glob(os.path.join(DATA, "*.fasta"))):
        name = os.path.basename(path)[:-6]
        a = Alignment()
        a.initByName("sequences", read_fasta(path))
        loci[name] = {"taxa": a.taxa, "patterns": ["".join("ACGT-"[s] for s in row) for row in a.patterns],
                      "weights": a.weights.tolist(), "sites": int(a.weights.sum())}
    out = {"source": "genomescale/starbeast2 tutorial/data/*.fasta (Canis, 16 loci); species = taxon name without its _a/_b suffix "
                     "(taxonsuperset, tutorial/CanisPhylogeny/Canis-Phylogeny.xml:378-411)",
           "loci": loci}
    json.dump(out, open(os.path.join(HERE, "canis_c1.json"), "w"), indent=0)
    print({k: (len(v["taxa"]), len(v["weights"]), v["sites"]) for k, v in loci.items()})


if __name__ == "__main__":
    main()
