"""CPU tier: --parse-fasta semantics (createIndexFromFASTA, GenbankTools.h:224-260) of slam.parse_fasta."""
import numpy as np


def slow_reference(paths):
    """Line-by-line restatement of the reference loop (the thing parse_fasta's block scanner must equal)."""
    entries = []
    for path in paths:
        data = open(path, "rb").read().replace(b"\r\n", b"\n").replace(b"\r", b"\n")
        bases, tag = b"", b""
        for line in data.split(b"\n"):
            if not line:
                continue
            if line[:1] == b">":
                if bases:
                    entries.append((bases, tag))
                bases, tag = b"", b""
                sp = line.find(b" ")
                if sp not in (-1, 0):
                    tag = line[1:sp]
            else:
                bases += line
        if bases:
            entries.append((bases, tag))
    return [(b.upper(), t) for b, t in entries]


def test_parse_fasta_matches_line_by_line_restatement(pkg, tmp_path):
    from kslam_b200 import slam
    blobs = [b">g0 first\nACGT\nacgtn\n>g1 second entry\nTTTT\n\n\nGG\n>nospace\nAAAA\n> leading\nCC\n>empty follows\n>g5 x\nGATTACA",
             b"ACGT\n>late header\nGG\n", b">a b\r\nAC\r\nGT\r\n>c d\rTT\rAA", b"", b">only header\n", b"\n\n>x y\n\nAC\n\n"]
    paths = []
    for k, b in enumerate(blobs):
        p = tmp_path / f"f{k}.fa"; p.write_bytes(b); paths.append(str(p))
    for sel in ([0], [1], [2], [3], [4], [5], list(range(len(blobs)))):
        ps = [paths[i] for i in sel]
        bases, offs, tags = slam.parse_fasta(ps)
        got = [(bytes(bases[int(offs[i]):int(offs[i + 1])]), tags[i]) for i in range(len(tags))]
        assert got == slow_reference(ps), sel
