"""Host-stage profile of a metagenomic batch without a GPU: config-2-like database (strains in a phylogeny, GenBank genes,
taxonomy), alignments from the reference (oracle/_ref, test infrastructure), then kslam_batch_outputs + kslam_taxa_results
timed with KSLAM_SAM_TRACE=1. Usage: python tests/perf/prof_meta_host.py [pairs] [threads]"""
import os
import pathlib
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import _lib as T  # noqa: E402
from test_taxon_host import make_db  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pkg = T.load_pkg()
tmp = pathlib.Path(tempfile.mkdtemp())
gb, go, _, _, _, _, taxdb, paths = make_db(pkg, tmp, n_strains=100, length=100_000, files=1)
ix = pkg.Index.parse_genbank(paths)
rb, ro, _ = pkg.synth.paired_reads(gb, go, pairs, seed=2)
cache = f"/tmp/prof_meta_{pairs}.npz"
if os.path.exists(cache):
    z = np.load(cache); ov, pool, pr = z["ov"], z["pool"], z["pr"]
else:
    R = T.Ref(gb, go, rb, ro, T.default_params(report_cigar=1))
    t0 = time.time(); R.align_to_database(); ov, pool, pr = R.screen_and_pair(); R.close()
    print(f"reference alignToDatabase + pairing: {time.time() - t0:.1f}s, {len(ov)} alignments, {len(pr)} pair records")
    np.savez(cache, ov=ov, pool=pool, pr=pr)
n = len(ro) - 1
quals = np.full(len(rb), ord("I"), np.uint8)
ids = [b"r%d" % (i % pairs) for i in range(n)]
idb = np.frombuffer(b"".join(ids), np.uint8); ido = np.zeros(n + 1, np.uint64); ido[1:] = np.cumsum([len(x) for x in ids])
os.environ["KSLAM_SAM_TRACE"] = "1"
w = pkg.SamWriter(index=ix, report_cigar=True, threads=threads)
db = pkg.TaxDb(taxdb)
for _ in range(2):
    taxa = pkg.Taxa()
    t0 = time.time()
    text, mi = w.batch(rb, ro, quals, ro, idb, ido, ov, pool, pr, taxdb=db, taxa=taxa)
    t1 = time.time()
    per_read, xml, abbr = taxa.results(db, pairs)
    t2 = time.time()
    print(f"kslam_batch_outputs: {t1 - t0:.3f}s for {pairs} pairs ({len(pr)} pair records), SAM {len(text) / 1e6:.1f} MB; "
          f"kslam_taxa_results: {t2 - t1:.3f}s, XML {len(xml) / 1e6:.1f} MB, {xml.count(b'<taxon>')} taxa")
