// slam_main.cpp — the `SLAM` executable: the reference's command line (/root/reference/src/main.cpp:24-169) and batch loop
// (metagenomicAnalysis_Low_Mem, /root/reference/src/SLAM.h:159-268) as a thin C++ host over the C ABI of libkslam.so.
//
//   SLAM [options] --db=DATABASE R1FILE [R2FILE]        align, SAM (--sam-file) and / or taxonomy XML (--output-file)
//   SLAM --parse-fasta   --output-file DB/database  a.fa b.fa ...
//   SLAM --parse-genbank --output-file DB/database  a.gbff ...
//   SLAM --parse-taxonomy --output-file DB/taxDB    names.dmp nodes.dmp
//
// Same options, defaults, files and messages as the reference (log.txt in the CWD included). Everything that computes is a
// library call: kslam_fastq_next (reader) | kslam_align_pair_batch / kslam_align_batch (GPU: alignToDatabase + score screen +
// getPairedOverlaps) | kslam_batch_outputs (host stages, SAM text, per-read taxa) run as a three-stage pipeline, one
// thread per stage, so batch i+2 is being read while batch i+1 is on the GPU and batch i is being written. There is no
// CPU fallback: without an sm_100 GPU the alignment modes stop with the library's error; the --parse-* modes are host-only.
// Option names may be abbreviated to an unambiguous prefix, as boost::program_options allows.
#include "../../include/kslam.h"
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Log {                                               // sequenceTools.h:150-179: "[t = 1.23s]\tmessage" lines in ./log.txt
  std::ofstream f;
  std::chrono::system_clock::time_point t0 = std::chrono::system_clock::now();
  std::mutex m;
  void write(const std::string &s) {
    std::lock_guard<std::mutex> g(m);
    if (!f.is_open()) { f.open("log.txt"); f.setf(std::ios::fixed, std::ios::floatfield); f.precision(2); }
    const double t = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now() - t0).count() / 1000.0;
    f << "[t = " << t << "s]\t" << s << std::endl;
  }
} g_log;
void log(const std::string &s) { g_log.write(s); }

struct Options {                                           // main.cpp:36-82, defaults included
  std::string db, outFileName, samFileName;
  uint32_t scoreThreshold = 0, match = 2, misMatch = 3, gapOpen = 5, gapExtend = 2;
  uint32_t numReads = UINT32_MAX, numReadsAtOnce = 10000000, numSAMAlignments = 10;
  double scoreFractionThreshold = 0.95;
  bool help = false, version = false, samXA = false, justAlign = false, noPseudoAssembly = false;
  bool parseGenbank = false, parseFasta = false, parseTaxonomy = false;
  std::vector<int> devices{0};                             // extension: --device N / --devices N,M,... (CUDA ordinals, one context each)
  uint32_t maxCigarOps = 0;                                // extension: --max-cigar-ops N: initial CIGAR pool stride (it grows when a CIGAR needs more)
  int partitionIndex = -1;                                 // extension: --partition-index 0|1 (-1 = decide from the database size)
  std::vector<std::string> inputs;
};

struct OptSpec { const char *name; int kind; };             // kind: 0 flag, 1 takes a value
const OptSpec kSpecs[] = {
    {"help", 0}, {"db", 1}, {"min-alignment-score", 1}, {"score-fraction-threshold", 1}, {"match-score", 1}, {"mismatch-penalty", 1},
    {"gap-open", 1}, {"gap-extend", 1}, {"num-reads", 1}, {"num-reads-at-once", 1}, {"output-file", 1}, {"sam-file", 1},
    {"num-alignments", 1}, {"sam-xa", 0}, {"version", 0}, {"just-align", 0}, {"no-pseudo-assembly", 0}, {"server", 0},
    {"input-file", 1}, {"parse-genbank", 0}, {"parse-fasta", 0}, {"parse-taxonomy", 0}, {"alignment-only", 0}, {"device", 1}, {"devices", 1}, {"partition-index", 1}, {"max-cigar-ops", 1}};

uint32_t to_u32(const std::string &name, const std::string &v) {
  size_t used = 0;
  unsigned long x = 0;
  try { x = std::stoul(v, &used); } catch (...) { used = 0; }
  if (used != v.size() || v.empty() || v[0] == '-' || x > UINT32_MAX) throw std::runtime_error("the argument ('" + v + "') for option '--" + name + "' is invalid");
  return (uint32_t)x;
}

Options parse_options(int argc, char **argv) {
  Options o;
  bool only_positional = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (only_positional || a.size() < 3 || a.compare(0, 2, "--") != 0) {
      if (a == "--") { only_positional = true; continue; }
      o.inputs.push_back(a);
      continue;
    }
    std::string name = a.substr(2), value;
    bool has_value = false;
    const size_t eq = name.find('=');
    if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); has_value = true; }
    const OptSpec *spec = nullptr;
    int matches = 0;
    for (const OptSpec &s : kSpecs) {
      if (name == s.name) { spec = &s; matches = 1; break; }
      if (std::string(s.name).compare(0, name.size(), name) == 0) { spec = &s; matches++; }
    }
    if (matches == 0) throw std::runtime_error("unrecognised option '--" + name + "'");
    if (matches > 1) throw std::runtime_error("option '--" + name + "' is ambiguous");
    name = spec->name;
    if (spec->kind == 1 && !has_value) {
      if (i + 1 >= argc) throw std::runtime_error("the required argument for option '--" + name + "' is missing");
      value = argv[++i];
    }
    if (spec->kind == 0 && has_value) throw std::runtime_error("option '--" + name + "' does not take any arguments");
    if (name == "help") o.help = true;
    else if (name == "version") o.version = true;
    else if (name == "db") o.db = value;
    else if (name == "min-alignment-score") o.scoreThreshold = to_u32(name, value);
    else if (name == "score-fraction-threshold") {
      size_t used = 0;
      try { o.scoreFractionThreshold = std::stod(value, &used); } catch (...) { used = 0; }
      if (used != value.size() || value.empty()) throw std::runtime_error("the argument ('" + value + "') for option '--" + name + "' is invalid");
    } else if (name == "match-score") o.match = to_u32(name, value);
    else if (name == "mismatch-penalty") o.misMatch = to_u32(name, value);
    else if (name == "gap-open") o.gapOpen = to_u32(name, value);
    else if (name == "gap-extend") o.gapExtend = to_u32(name, value);
    else if (name == "num-reads") o.numReads = to_u32(name, value);
    else if (name == "num-reads-at-once") o.numReadsAtOnce = to_u32(name, value);
    else if (name == "output-file") o.outFileName = value;
    else if (name == "sam-file") o.samFileName = value;
    else if (name == "num-alignments") o.numSAMAlignments = to_u32(name, value);
    else if (name == "sam-xa") o.samXA = true;
    else if (name == "just-align") o.justAlign = true;
    else if (name == "no-pseudo-assembly") o.noPseudoAssembly = true;
    else if (name == "input-file") o.inputs.push_back(value);
    else if (name == "parse-genbank") o.parseGenbank = true;
    else if (name == "parse-fasta") o.parseFasta = true;
    else if (name == "parse-taxonomy") o.parseTaxonomy = true;
    else if (name == "device") o.devices.assign(1, (int)to_u32(name, value));
    else if (name == "partition-index") o.partitionIndex = to_u32(name, value) ? 1 : 0;
    else if (name == "max-cigar-ops") o.maxCigarOps = to_u32(name, value);
    else if (name == "devices") {
      o.devices.clear();
      for (size_t at = 0; at <= value.size();) {
        const size_t comma = std::min(value.find(',', at), value.size());
        o.devices.push_back((int)to_u32(name, value.substr(at, comma - at)));
        at = comma + 1;
      }
    }
    // --server and --alignment-only are accepted and ignored, as in the reference (main.cpp never reads them)
  }
  return o;
}

void usage() {                                             // main.cpp:96-107
  std::cout << "Usage\tSLAM [option] --db=DATABASE R1FILE R2FILE\n"
               "\tAlign paired reads from R1FILE and R2FILE against DATABASE and perform metagenomic analysis\n"
               "or\tSLAM [option] --db=DATABASE R1FILE\n"
               "\tAlign reads from R1FILE against DATABASE and perform metagenomic analysis\n"
               "Allowed options:\n"
               "  --help                                produce help message\n"
               "  --db arg                              SLAM database directory which reads will be aligned against\n"
               "  --min-alignment-score arg (=0)        alignment score cutoff\n"
               "  --score-fraction-threshold arg (=0.95) screen alignments with scores < this*top score\n"
               "  --match-score arg (=2)                match score\n"
               "  --mismatch-penalty arg (=3)           mismatch penalty (positive)\n"
               "  --gap-open arg (=5)                   gap opening penalty (positive)\n"
               "  --gap-extend arg (=2)                 gap extend penalty (positive)\n"
               "  --num-reads arg (=4294967295)         Number of reads from R1/R2 File to align\n"
               "  --num-reads-at-once arg (=10000000)   Reduce RAM usage by only analysing \"arg\" reads at once, this will increase execution time\n"
               "  --output-file arg                     write to this file instead of stdout\n"
               "  --sam-file arg                        write SAM output to this file\n"
               "  --num-alignments arg (=10)            Number of alignments to report in SAM file\n"
               "  --sam-xa                              only output primary alignment lines, use XA field for secondary alignments\n"
               "  --version                             print version number\n"
               "  --just-align                          only perform alignments, not metagenomics\n"
               "  --no-pseudo-assembly                  do not link alignments together\n"
               "  --device arg (=0)                     CUDA device ordinal (this implementation)\n"
               "  --devices arg                         comma-separated CUDA ordinals: every batch is split into that many contiguous\n"
               "                                        ranges of read pairs, one context per entry (this implementation)\n"
               "  --partition-index arg                 1: range-partition the genome k-mer index by k-mer prefix over --devices and route\n"
               "                                        read k-mers / matches with NCCL all-to-alls; 0: replicate it; default: partition\n"
               "                                        when a replica would not fit a device (this implementation)\n"
               "  --max-cigar-ops arg (=32)             initial per-alignment CIGAR capacity; grown automatically (this implementation)\n\n";
}

// bounded hand-over between two pipeline stages
template <class T> struct Slot {
  std::mutex m; std::condition_variable cv; bool full = false; T item{};
  void put(T v) { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return !full; }); item = std::move(v); full = true; cv.notify_all(); }
  T take() { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return full; }); T v = std::move(item); full = false; cv.notify_all(); return v; }
};

struct Batch {                                             // one batch on its way through the pipeline; n_reads == 0 ends the run
  kslam_read_batch reads{};
  std::vector<kslam_overlap> overlaps; std::vector<uint32_t> cigars; std::vector<kslam_pair> pairs;   // copies: the ctx reuses its buffers
  std::vector<kslam_pair_compact> compact; std::vector<kslam_far_mates> far; uint32_t limit = 0;       // runs without --sam-file (paired)
  std::string error;
};

bool write_file(const std::string &path, const char *text, uint64_t len) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = fwrite(text, 1, len, f) == len;
  return fclose(f) == 0 && ok;
}

// Stage 2 for one batch. The batch is cut into contiguous ranges of read pairs that keep mates together (R1 i and R2
// i + mid), one per context (one range = the whole batch with a single context); every context runs the whole path on its
// range — alignToDatabase over a replicated index, or the collective kslam_comm_align_resident over a k-mer-range
// partitioned one — and the results are joined in range order with read / overlap / cigar / pair indices rebased, which
// is the order one context would have produced (pairing is per pair, seeds are per (read, genome); k-slam_b200/shard.py
// states the same rules). `compact` (paired runs without --sam-file): only the 24-byte pair records and the mates of the
// pairs beyond the batch's insert-size limit leave the GPUs (kslam_fetch_pairs_compact).
void align_batch(const std::vector<kslam_ctx *> &ctxs, const std::vector<kslam_comm *> &comms, bool isPaired, bool compact, Batch *b) {
  const kslam_read_batch &r = b->reads;
  const bool partitioned = !comms.empty();                 // every rank takes part in every batch, even with an empty range
  // (a paired batch with an odd read count — the reference's R1/R2 size check lets n2 = n1 + 1 through, FASTQsequence.h:118-122 —
  // pairs its last read by index modulo the midpoint, which no contiguous split reproduces: such a batch runs on one context)
  const bool odd = isPaired && (r.n_reads & 1);
  if (partitioned && odd) { b->error = "a paired batch with an odd read count cannot be sharded over a partitioned index"; return; }
  const size_t G = odd ? 1 : ctxs.size();
  struct Shard {
    uint64_t lo = 0, hi = 0; std::vector<char> bases; std::vector<uint64_t> offs;
    kslam_pairs p{}; kslam_alignments a{}; kslam_pairs_compact c{}; std::string error;
  };
  std::vector<Shard> shards(G);
  const uint64_t units = isPaired ? r.n_reads / 2 : r.n_reads, mid = r.n_reads / 2;
  auto run_shard = [&](size_t g) {
    Shard &s = shards[g];
    const uint64_t cnt = s.hi - s.lo;
    const char *base_ptr = r.bases;
    const uint64_t *offs_ptr = r.offs;
    uint64_t n_here = r.n_reads;
    if (G > 1 || partitioned) {
      if (isPaired) {                                        // R1 block of the range, then its R2 block, in one array
        const uint64_t a0 = r.offs[s.lo], a1 = r.offs[s.hi], b0 = r.offs[mid + s.lo], b1 = r.offs[mid + s.hi];
        s.bases.resize((a1 - a0) + (b1 - b0) + 1);
        memcpy(s.bases.data(), r.bases + a0, a1 - a0);
        memcpy(s.bases.data() + (a1 - a0), r.bases + b0, b1 - b0);
        s.offs.assign(2 * cnt + 1, 0);
        for (uint64_t i = 0; i <= cnt; i++) s.offs[i] = r.offs[s.lo + i] - a0;
        for (uint64_t i = 1; i <= cnt; i++) s.offs[cnt + i] = (a1 - a0) + (r.offs[mid + s.lo + i] - b0);
        base_ptr = s.bases.data(); n_here = 2 * cnt;
      } else {                                               // a contiguous slice of the batch: only the offsets are rebased
        s.offs.assign(cnt + 1, 0);
        for (uint64_t i = 0; i <= cnt; i++) s.offs[i] = r.offs[s.lo + i] - r.offs[s.lo];
        base_ptr = r.bases + r.offs[s.lo]; n_here = cnt;
      }
      offs_ptr = s.offs.data();
    }
    kslam_ctx *ctx = ctxs[g];
    int rc = kslam_upload_reads(ctx, n_here, base_ptr, offs_ptr);
    if (rc != KSLAM_OK && partitioned) {                     // the other ranks are about to enter the collective: tell them
      s.error = kslam_last_error(ctx);
      kslam_comm_abort_batch(comms[g]);
      return;
    }
    const int fetch_alignments = isPaired ? 0 : 1;
    if (rc == KSLAM_OK) rc = partitioned ? kslam_comm_align_resident(comms[g], fetch_alignments, isPaired ? nullptr : &s.a)
                                         : kslam_align_resident(ctx, fetch_alignments, isPaired ? nullptr : &s.a);
    if (rc == KSLAM_OK && isPaired) rc = kslam_pair_batch(ctx, compact ? 0 : 1, compact ? nullptr : &s.p);
    if (rc == KSLAM_OK && isPaired && compact) rc = kslam_fetch_pairs_compact(ctx, 0, &s.c);
    if (rc != KSLAM_OK) s.error = kslam_last_error(ctx);
  };
  std::vector<std::thread> th;
  for (size_t g = 0; g < G; g++) {
    shards[g].lo = units * g / G; shards[g].hi = units * (g + 1) / G;
    if (shards[g].hi == shards[g].lo && !partitioned && G > 1) continue;
    if (G == 1) run_shard(0); else th.emplace_back(run_shard, g);
  }
  for (auto &t : th) t.join();
  for (Shard &s : shards) if (!s.error.empty()) { b->error = s.error; return; }
  if (compact && isPaired) {
    // compact records: rebase the read-pair ids; ONE insert-size limit for the batch (getMaxAllowedInsertSize is a statistic
    // of the whole batch, PairedOverlap.h:314-360), so with several ranges the far-mates tables are fetched again for it
    std::vector<uint64_t> pair_base(G, 0);
    for (size_t g = 0; g < G; g++) {
      const Shard &s = shards[g];
      pair_base[g] = b->compact.size();
      b->compact.insert(b->compact.end(), s.c.pairs, s.c.pairs + s.c.n_pairs);
      for (uint64_t i = pair_base[g]; i < b->compact.size(); i++) b->compact[i].pair_id += (uint32_t)s.lo;
    }
    b->limit = G == 1 ? shards[0].c.insert_size_limit : kslam_insert_size_limit_compact(b->compact.data(), b->compact.size(), 0);
    for (size_t g = 0; g < G; g++) {
      Shard &s = shards[g];
      uint64_t n_far = s.c.n_far; const kslam_far_mates *far = s.c.far;
      if (s.c.insert_size_limit != b->limit && kslam_fetch_far_mates(ctxs[g], b->limit, &n_far, &far) != KSLAM_OK) { b->error = kslam_last_error(ctxs[g]); return; }
      const uint64_t at = b->far.size();
      b->far.insert(b->far.end(), far, far + n_far);
      for (uint64_t i = at; i < b->far.size(); i++) b->far[i].pair_index += (uint32_t)pair_base[g];
    }
    return;
  }
  for (Shard &s : shards) {
    const uint64_t cnt = s.hi - s.lo, ov_base = b->overlaps.size(), cg_base = b->cigars.size();
    const kslam_overlap *ov = isPaired ? s.p.sorted_overlaps : s.a.overlaps;
    const uint64_t n_ov = isPaired ? s.p.n_sorted : s.a.n_overlaps;
    const uint32_t *cg = isPaired ? s.p.cigar_pool : s.a.cigar_pool;
    const uint64_t n_cg = isPaired ? s.p.n_cigar_words : s.a.n_cigar_words;
    if (n_ov) b->overlaps.insert(b->overlaps.end(), ov, ov + n_ov);
    if (G > 1 || partitioned)
      for (uint64_t i = ov_base; i < b->overlaps.size(); i++) {
        kslam_overlap &o = b->overlaps[i];
        o.read = isPaired ? (uint32_t)(o.read < cnt ? s.lo + o.read : mid + s.lo + (o.read - cnt)) : (uint32_t)(o.read + s.lo);
        if (o.cigar_len) o.cigar_off += (uint32_t)cg_base;
      }
    if (n_cg) b->cigars.insert(b->cigars.end(), cg, cg + n_cg);
    if (isPaired) {
      const uint64_t pr_base = b->pairs.size();
      if (s.p.n_pairs) b->pairs.insert(b->pairs.end(), s.p.pairs, s.p.pairs + s.p.n_pairs);
      if (ov_base)
        for (uint64_t i = pr_base; i < b->pairs.size(); i++) {
          if (b->pairs[i].r1_idx >= 0) b->pairs[i].r1_idx += (int32_t)ov_base;
          if (b->pairs[i].r2_idx >= 0) b->pairs[i].r2_idx += (int32_t)ov_base;
        }
    }
  }
}

int run_alignment(const Options &o, const std::string &commandLine) {      // metagenomicAnalysis_Low_Mem, SLAM.h:159-268
  log("Performing metagenomic analysis");
  const bool isPaired = o.inputs.size() == 2;
  const bool wantSam = !o.samFileName.empty();
  kslam_params prm;
  memset(&prm, 0, sizeof prm);
  prm.match = (uint8_t)o.match; prm.mismatch = (uint8_t)o.misMatch; prm.gap_open = (uint8_t)o.gapOpen; prm.gap_extend = (uint8_t)o.gapExtend;   // ssw_cpp.cpp:114-117
  // sw_score is 16 bits (ssw_cpp.h:16): a larger threshold screens everything, as in the reference
  prm.score_threshold = (uint16_t)(o.scoreThreshold > 65535u ? 65535u : o.scoreThreshold);
  prm.report_cigar = wantSam ? 1 : 0; prm.device = o.devices[0]; prm.max_cigar_ops = o.maxCigarOps;
  if (!kslam_params_fast(&prm))
    log("Scoring parameters outside gap-extend < gap-open, mismatch <= 2 * gap-extend: Smith-Waterman runs the lane-for-lane restatement of SSW's striped kernels (same results, slower)");
  // CUDA context creation (0.3-1.5 s) runs while the taxonomy and the database are read from disk
  // one context per entry of --devices (read pairs shard trivially; the index is replicated or partitioned below)
  const size_t G = o.devices.size();
  std::vector<kslam_ctx *> ctxs(G, nullptr);
  std::string create_error;
  std::thread create([&] {
    kslam_params p = prm;
    for (size_t g = 0; g < G; g++) {
      p.device = o.devices[g];
      if (kslam_create(&p, &ctxs[g]) != KSLAM_OK) { create_error = kslam_last_error(nullptr); return; }
    }
  });
  struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } join_on_any_return{create};
  kslam_taxdb *taxdb = nullptr;
  kslam_taxa *taxa = nullptr;
  if (!o.justAlign) {
    log("Building taxonomy index");
    if (kslam_taxdb_open((o.db + "/taxDB").c_str(), &taxdb) != KSLAM_OK) { std::cerr << "SLAM: unable to open taxonomy index file " << o.db << "/taxDB\n"; return 2; }
    log("Built a taxonomy tree with " + std::to_string(kslam_taxdb_size(taxdb)) + " nodes");
    kslam_taxa_create(&taxa);
  }
  log("Building index from serial file\t " + o.db + "/database");
  kslam_index *index = nullptr;
  if (kslam_index_read((o.db + "/database").c_str(), &index) != KSLAM_OK) { std::cerr << "SLAM: " << kslam_index_error() << " (" << o.db << "/database)\n"; return 2; }
  kslam_sam_db db;
  kslam_index_db(index, &db);

  create.join();
  if (!create_error.empty()) { std::cerr << "SLAM: " << create_error << "\n"; return 3; }
  log("Getting k-mers from index");
  // Replicate the genome k-mer index in every context, or — when a replica (16 B per record resident, three times that
  // while it is built) would not fit a device, or on request — range-partition it by k-mer prefix over the devices
  // (kslam_load_genomes_part + kslam_comm: read k-mers are routed to their key owner by NCCL all-to-all).
  std::vector<kslam_comm *> comms;
  {
    uint64_t n_records = 0;
    for (uint64_t e = 0; e < db.n_entries; e++) { const uint64_t len = db.offs[e + 1] - db.offs[e]; if (len >= 32) n_records += (len - 32) / 16 + 1; }
    uint64_t free_b = 0, total_b = 0;
    kslam_device_memory(o.devices[0], &free_b, &total_b);
    const bool too_big = (double)n_records * 48.0 + (double)db.offs[db.n_entries] * 1.8 > 0.8 * (double)total_b;
    const bool partition = G > 1 && (o.partitionIndex == 1 || (o.partitionIndex < 0 && too_big));
    if (too_big && !partition) log("Warning: the genome k-mer index may not fit one device; give several --devices to range-partition it");
    std::vector<int> rc(G, 0);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++)
      th.emplace_back([&, g] { rc[g] = partition ? kslam_load_genomes_part(ctxs[g], db.n_entries, db.bases, db.offs, (uint32_t)g, (uint32_t)G)
                                                 : kslam_load_genomes(ctxs[g], db.n_entries, db.bases, db.offs); });
    for (auto &t : th) t.join();
    for (size_t g = 0; g < G; g++) if (rc[g] != KSLAM_OK) { std::cerr << "SLAM: " << kslam_last_error(ctxs[g]) << "\n"; return 3; }
    if (partition) {
      log("Genome k-mer index range-partitioned over " + std::to_string(G) + " devices (" + std::to_string(n_records) + " records)");
      comms.assign(G, nullptr);
      if (kslam_comm_init_all((uint32_t)G, ctxs.data(), comms.data()) != KSLAM_OK) { std::cerr << "SLAM: " << kslam_last_error(ctxs[0]) << "\n"; return 3; }
    }
  }

  kslam_fastq *reader = nullptr;
  if (kslam_fastq_open(o.inputs[0].c_str(), isPaired ? o.inputs[1].c_str() : nullptr, 0, &reader) != KSLAM_OK) {
    log("FASTQ file " + o.inputs[0] + " bad");
    std::cerr << "SLAM: " << kslam_last_error(nullptr) << "\n";
    return 2;
  }
  kslam_fastq_set_ring(reader, 6);                         // one set being filled, one in each hand-over slot, one in each later stage, one spare

  FILE *sam = nullptr;
  if (wantSam) {
    sam = fopen(o.samFileName.c_str(), "wb");
    if (!sam) { std::cerr << "SLAM: unable to open " << o.samFileName << "\n"; return 2; }
    char *hdr = nullptr; uint64_t n = 0;
    kslam_sam_header(&db, commandLine.c_str(), &hdr, &n);
    fwrite(hdr, 1, n, sam);
    kslam_sam_free(hdr);
  }

  Slot<Batch *> to_gpu, to_host;
  std::thread ingest([&] {                                 // stage 1: FASTQ reader (SLAM.h:194-208)
    uint64_t numReads = 0;
    const bool lowMem = o.numReadsAtOnce != UINT32_MAX;    // main.cpp:146-166: otherwise metagenomicAnalysis reads everything in one go
    for (uint64_t go = 0;; go++) {
      Batch *b = new Batch();
      uint64_t perGo = lowMem ? o.numReadsAtOnce : o.numReads;
      if (numReads + perGo > o.numReads) perGo = o.numReads - numReads;
      if (!lowMem && go > 0) perGo = 0;
      if (isPaired) log("Getting reads from FASTQ files " + o.inputs[0] + " and " + o.inputs[1]);
      else log("Getting reads from FASTQ file " + o.inputs[0]);
      if (perGo == 0) b->reads.n_reads = 0;
      else if (kslam_fastq_next(reader, perGo, &b->reads) != KSLAM_OK) { b->error = kslam_fastq_error(reader); b->reads.n_reads = 0; }
      numReads += isPaired ? b->reads.n_reads / 2 : b->reads.n_reads;
      const bool last = b->reads.n_reads == 0;
      to_gpu.put(b);
      if (last) return;
    }
  });
  std::thread gpu([&] {                                    // stage 2: alignToDatabase + score screen + getPairedOverlaps on the GPU(s)
    for (;;) {
      Batch *b = to_gpu.take();
      if (b->reads.n_reads && b->error.empty()) align_batch(ctxs, comms, isPaired, isPaired && !wantSam, b);
      const bool last = b->reads.n_reads == 0 || !b->error.empty();
      to_host.put(b);
      if (last) return;
    }
  });

  kslam_sam_params sp;
  memset(&sp, 0, sizeof sp);
  sp.num_alignments = o.numSAMAlignments; sp.pseudo_assembly = o.noPseudoAssembly ? 0 : 1; sp.report_cigar = wantSam ? 1 : 0;
  sp.sam_xa = o.samXA ? 1 : 0; sp.threads = 0; sp.score_fraction_threshold = o.scoreFractionThreshold;
  uint64_t numReads = 0;
  std::string error;
  for (;;) {                                               // stage 3: host stages, SAM text, per-read taxa (SLAM.h:215-250)
    Batch *b = to_host.take();
    if (!b->error.empty()) error = b->error;
    if (b->reads.n_reads == 0 || !error.empty()) { delete b; break; }
    numReads += isPaired ? b->reads.n_reads / 2 : b->reads.n_reads;
    char *text = nullptr; uint64_t len = 0;
    int rc;
    if (wantSam) log("Writing SAM output");
    if (isPaired && !wantSam) {
      kslam_pairs_compact pc;
      pc.n_pairs = b->compact.size(); pc.pairs = b->compact.data(); pc.insert_size_limit = b->limit; pc.n_far = b->far.size(); pc.far = b->far.data();
      rc = kslam_batch_outputs_compact(&sp, &db, &b->reads, &pc, nullptr, taxdb, taxa);
    } else if (isPaired) {
      kslam_pairs p;
      p.n_sorted = b->overlaps.size(); p.sorted_overlaps = b->overlaps.data(); p.n_cigar_words = b->cigars.size(); p.cigar_pool = b->cigars.data();
      p.n_pairs = b->pairs.size(); p.pairs = b->pairs.data();
      rc = kslam_batch_outputs(&sp, &db, &b->reads, &p, wantSam, &text, &len, nullptr, taxdb, taxa);
    } else {
      kslam_alignments a;
      a.n_overlaps = b->overlaps.size(); a.overlaps = b->overlaps.data(); a.n_cigar_words = b->cigars.size(); a.cigar_pool = b->cigars.data();
      rc = kslam_batch_outputs_single(&sp, &db, &b->reads, &a, o.scoreThreshold, wantSam, &text, &len, taxdb, taxa);
    }
    if (rc != KSLAM_OK) error = "host stages failed (" + std::to_string(rc) + ")";
    if (text) { if (sam && fwrite(text, 1, len, sam) != len) error = "short write to " + o.samFileName; kslam_sam_free(text); }
    if (!o.justAlign) log("Processed\t" + std::to_string(numReads) + "\t reads");
    delete b;
    if (!error.empty()) break;
  }
  if (!error.empty()) {                                    // like the reference's uncaught exception: stop here, upstream stages die with the process
    std::cerr << "SLAM: " << error << std::endl;
    if (sam) fclose(sam);
    _Exit(4);
  }
  ingest.join(); gpu.join();
  if (sam) fclose(sam);
  int rc = 0;
  if (!o.justAlign) {                                      // SLAM.h:256-265
    char *per_read = nullptr, *xml = nullptr, *abbreviated = nullptr;
    uint64_t n1 = 0, n2 = 0, n3 = 0;
    log("Writing per read results");
    log("Combining taxonomies");
    if (kslam_taxa_results(taxa, taxdb, (uint32_t)numReads, &per_read, &n1, &xml, &n2, &abbreviated, &n3) != KSLAM_OK) { std::cerr << "SLAM: writing the results failed\n"; rc = 4; }
    else {
      log("Writing results file");
      // SLAM.h:142 vs :256 — the all-at-once variant names the per-read file without the underscore
      if (!write_file(o.outFileName + (o.numReadsAtOnce != UINT32_MAX ? "_PerRead" : "PerRead"), per_read, n1)) rc = 4;
      if (o.outFileName.size()) {
        if (!write_file(o.outFileName, xml, n2) || !write_file(o.outFileName + "_abbreviated", abbreviated, n3)) rc = 4;
      } else fwrite(xml, 1, n2, stdout);
      if (rc) std::cerr << "SLAM: unable to write the result files\n";
    }
    kslam_sam_free(per_read); kslam_sam_free(xml); kslam_sam_free(abbreviated);
  }
  log("Done");
  // Every output is written and closed. KSLAM_FAST_EXIT=1 leaves without unmapping the pinned buffers and tearing the CUDA
  // contexts down (0.5 s of a 3.3 s run) — measured: the driver then cleans up behind the dead process and the NEXT
  // process pays 1.6 s in its context creation, so back-to-back runs get slower; off by default.
  fflush(stdout);
  if (getenv("KSLAM_FAST_EXIT")) _Exit(rc);
  kslam_fastq_close(reader);
  for (kslam_comm *m : comms) kslam_comm_destroy(m);
  for (kslam_ctx *c : ctxs) kslam_destroy(c);
  kslam_taxa_destroy(taxa);
  kslam_taxdb_close(taxdb);
  kslam_index_free(index);
  return rc;
}

}  // namespace

int main(int argc, char **argv) {
  std::string commandLine;                                 // main.cpp:25-31, goes into the SAM @PG line
  for (int i = 0; i < argc; i++) { if (i) commandLine += " "; commandLine += argv[i]; }
  Options o;
  try { o = parse_options(argc, argv); } catch (const std::exception &e) { std::cerr << "SLAM: " << e.what() << "\n"; return 2; }
  if (o.version) { std::cout << "1.0" << std::endl; return 1; }
  if (o.help || argc == 1) { usage(); return 1; }
  std::vector<const char *> paths;
  for (auto &s : o.inputs) paths.push_back(s.c_str());
  if (o.parseGenbank || o.parseFasta) {                    // main.cpp:108-119
    log(o.parseGenbank ? "Parsing Genbank" : "Parsing FASTA");
    for (auto &s : o.inputs) log("Parsing\t" + s);
    kslam_index *index = nullptr;
    int rc = o.parseGenbank ? kslam_index_parse_genbank(paths.data(), paths.size(), &index) : kslam_index_parse_fasta(paths.data(), paths.size(), &index);
    if (rc == KSLAM_OK) rc = kslam_index_write(index, o.outFileName.c_str());
    if (rc != KSLAM_OK) { std::cerr << "SLAM: " << kslam_index_error() << "\n"; return 2; }
    kslam_index_free(index);
    return 0;
  }
  if (o.parseTaxonomy) {                                   // main.cpp:120-130
    log("Parsing taxonomy");
    if (o.inputs.size() != 2) { std::cout << "Provide names.dmp and nodes.dmp\n"; return 1; }
    if (kslam_taxdb_build(o.inputs[0].c_str(), o.inputs[1].c_str(), o.outFileName.c_str()) != KSLAM_OK) { std::cerr << "SLAM: unable to build the taxonomy index\n"; return 2; }
    return 0;
  }
  if (o.inputs.size() == 1 || o.inputs.size() == 2) {
    if (o.db.empty()) { std::cerr << "SLAM: --db is required\n"; return 2; }
    return run_alignment(o, commandLine);
  }
  return 0;
}
