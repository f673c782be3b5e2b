"""Reading alignments the way the StarBEAST2 tutorials ship them (FASTA; tutorial/data/*.fasta) -- the step before the hot
path.  Pattern compression itself is plugins.Alignment (BEAST's Alignment.calcPatterns)."""
from __future__ import annotations

from typing import Dict


def read_fasta(path: str) -> Dict[str, str]:
    """{taxon: sequence}; the taxon is the first word of the header line."""
    seqs: Dict[str, str] = {}
    name = None
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                name = line[1:].split()[0]
                seqs[name] = ""
            elif name is not None:
                seqs[name] += line
    return seqs
