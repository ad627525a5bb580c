// taxon.cu — taxonomy database, lowest common ancestor, per-read taxon assignment and the metagenomic result files
// (host code; the part of the reference's batch loop after the SAM records, /root/reference/src/SLAM.h:243-265).
//
// north_star keeps LCA assignment and output formatting on the host; they are restated here so the library carries a
// run from FASTQ to the XML / _PerRead / _abbreviated files. Restated functions:
//   TaxonomyDB: readTaxonomyIndex, parseNodesDump, parseNamesDump, writeTaxonomyIndex, getParentTaxID,
//               getLowestCommonAncestor, getLineage, getScientificName, getRank     TaxonomyDatabase.h:95-265
//   getResultFromPairedOverlaps, convertAlignmentsToIdentifiedTaxonomies_parallel        MetagenomicResults.h:88-111,182-197
//   combineTaxonomies, combineRangeOfIdentifiedTaxonomy                                  MetagenomicResults.h:117-176
//   sortResults, writeResults, getXML, correctXML, writeAbbreviatedResultsFile, writePerReadResults   :213-369,455-463
//   Gene::operator==, geneSort                                                           GenbankTools.h:82-89,116-125
// The reference decides ties with std::sort; the same std::sort calls are issued here over sequences that compare the
// same way, so the permutations coincide. One exception is documented in include/kslam.h: combineTaxonomies uses
// __gnu_parallel::sort, whose order among reads of one taxon depends on the OpenMP thread count; this file uses the
// sequential std::sort (the reference with one thread). That order only shows when two genes compare equal
// (same protein id and product) but differ in their other fields.
#include "common.cuh"
#include "host_stages.h"
#include <algorithm>
#include <atomic>
#include <fstream>
#include <stdexcept>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unordered_map>

using namespace kslam_host;

namespace {

struct TaxEntry {                                          // TaxonomyEntry, TaxonomyDatabase.h:24-43 (fields that are read)
  uint32_t taxonomyID = 0, parentTaxonomyID = 0;
  std::string scientificName, rank;
};

// The dump files (nodes.dmp / names.dmp) are read whole and walked as views: no per-line string, no token vector.
struct DumpFile {
  std::string text;
  explicit DumpFile(const char *path, const char *what) {
    FILE *f = fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("unable to open ") + what + " file");
    char buf[1 << 16];
    for (size_t got; (got = fread(buf, 1, sizeof buf, f)) > 0;) text.append(buf, got);
    fclose(f);
  }
  // calls fn(begin, end) for every '\n'-terminated line and for what follows the last '\n' (possibly empty): the lines the
  // reference's `while (good()) getline` loop sees (TaxonomyDatabase.h:99-101)
  template <class F> void for_each_line(F fn) const {
    const char *p = text.data(), *end = p + text.size();
    for (;;) {
      const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
      fn(p, nl ? nl : end);
      if (!nl) break;
      p = nl + 1;
    }
  }
};
struct Field { const char *b, *e; size_t size() const { return (size_t)(e - b); } };
// The maximal runs of non-separator bytes of [b, e), at most cap of them; returns how many there are in total (what the
// reference's tokenise, sequenceTools.h:117-133, would return as tokens.size())
template <class IsSep> size_t split_fields(const char *b, const char *e, IsSep is_sep, Field *out, size_t cap) {
  size_t n = 0;
  while (b < e) {
    while (b < e && is_sep(*b)) b++;
    if (b == e) break;
    const char *s = b;
    while (b < e && !is_sep(*b)) b++;
    if (n < cap) out[n] = Field{s, b};
    n++;
  }
  return n;
}
// std::stoi of a field (same acceptance and the same exceptions as the reference's stoi(tokens[i]))
int field_int(const Field &f) { return std::stoi(std::string(f.b, f.e)); }

}  // namespace

struct kslam_taxdb {
  // the same container as the reference's taxIDsAndEntries filled by the same insert sequence, so that the iteration
  // order writeTaxonomyIndex depends on is the same
  std::unordered_map<uint32_t, TaxEntry> nodes;

  // Flat copy of the parent links for the LCA walks (built once by finalize()): node i has id[i]; next_id[i] is what
  // getParentTaxID answers for it (0 ends the walk: the parent is the root, id 1, or absent); next_idx[i] is the node that
  // id refers to, or -1 when it is not in the database (the walk records the id and ends). An open-addressing table maps
  // an id to its node. A walk is then a chain of 4-byte loads in a dense array instead of one hash-map probe per level —
  // with the NCBI taxonomy (2.5 M nodes, ~30 levels) the probes were the cost of the whole taxonomy stage.
  std::vector<uint32_t> flat_id, flat_next_id, table;
  std::vector<int32_t> flat_next_idx;
  uint32_t table_mask = 0;
  static uint32_t hash(uint32_t id) { return id * 2654435761u; }
  int32_t find(uint32_t id) const {
    if (table.empty()) return -1;
    for (uint32_t h = hash(id) & table_mask;; h = (h + 1) & table_mask) {
      const uint32_t slot = table[h];
      if (slot == 0) return -1;
      if (flat_id[slot - 1] == id) return (int32_t)(slot - 1);
    }
  }
  void finalize() {
    const size_t n = nodes.size();
    flat_id.clear(); flat_next_id.clear(); flat_next_idx.assign(n, -1);
    flat_id.reserve(n); flat_next_id.reserve(n);
    size_t cap = 16;
    while (cap < 2 * n + 2) cap <<= 1;
    table.assign(cap, 0); table_mask = (uint32_t)(cap - 1);
    for (auto &e : nodes) {
      flat_id.push_back(e.first);
      flat_next_id.push_back(e.second.parentTaxonomyID != 1 ? e.second.parentTaxonomyID : 0);
      uint32_t h = hash(e.first) & table_mask;
      while (table[h]) h = (h + 1) & table_mask;
      table[h] = (uint32_t)flat_id.size();
    }
    for (size_t i = 0; i < n; i++) flat_next_idx[i] = flat_next_id[i] ? find(flat_next_id[i]) : -1;
  }
  // the ids from taxID up to the end of its walk, leaf first, into out[cap]; returns the count, or -1 when cap is too small
  int path_of(uint32_t taxID, uint32_t *out, int cap) const {
    int len = 0;
    if (taxID == 0) return 0;
    int32_t idx = find(taxID);
    out[len++] = taxID;
    while (idx >= 0) {
      const uint32_t nid = flat_next_id[idx];
      if (nid == 0) break;
      if (len >= cap) return -1;
      out[len++] = nid;
      idx = flat_next_idx[idx];
    }
    return len;
  }

  void read_index(const char *path) {                      // readTaxonomyIndex, TaxonomyDatabase.h:166-183
    std::ifstream in(path);
    if (!in.is_open()) throw std::runtime_error("unable to open taxonomy index file");
    for (std::string line; getline(in, line);) {
      TaxEntry e;
      e.taxonomyID = stoi(line);
      getline(in, line); e.parentTaxonomyID = stoi(line);
      getline(in, line); e.scientificName = line;
      getline(in, line); e.rank = line;
      nodes.insert({e.taxonomyID, e});
    }
  }
  // parseNodesDump, TaxonomyDatabase.h:95-117: fields are separated by any run of tabs and bars; a line with at least three
  // of them is a node (id, parent id, rank). A repeated id keeps its first parent and rank (:111-114).
  void parse_nodes(const char *path) {
    const DumpFile dump(path, "nodes");
    dump.for_each_line([&](const char *b, const char *e) {
      Field f[3];
      if (split_fields(b, e, [](char c) { return c == '\t' || c == '|'; }, f, 3) < 3) return;
      const uint32_t id = (uint32_t)field_int(f[0]), up = (uint32_t)field_int(f[1]);
      auto slot = nodes.find(id);
      if (slot != nodes.end()) return;
      TaxEntry &node = nodes[id];
      node.taxonomyID = id; node.parentTaxonomyID = up; node.rank.assign(f[2].b, f[2].e);
    });
  }
  // parseNamesDump, :118-151: fields are separated by bars, each loses one leading and then one trailing tab when it is
  // longer than one byte; a line with at least four fields whose fourth is "scientific name" names the node of field one
  // (the id is converted before the class is looked at, as in the reference: a malformed id throws either way).
  void parse_names(const char *path) {
    const DumpFile dump(path, "names");
    static const char kClass[] = "scientific name";
    dump.for_each_line([&](const char *b, const char *e) {
      Field f[4];
      if (split_fields(b, e, [](char c) { return c == '|'; }, f, 4) < 4) return;
      for (Field &x : f)
        if (x.size() > 1) {
          if (*x.b == '\t') x.b++;
          if (x.e[-1] == '\t') x.e--;
        }
      const uint32_t id = (uint32_t)field_int(f[0]);
      if (f[3].size() != sizeof kClass - 1 || memcmp(f[3].b, kClass, sizeof kClass - 1) != 0) return;
      auto slot = nodes.find(id);
      if (slot == nodes.end()) { TaxEntry &node = nodes[id]; node.taxonomyID = id; node.scientificName.assign(f[1].b, f[1].e); }
      else slot->second.scientificName.assign(f[1].b, f[1].e);
    });
  }
  uint32_t parent(uint32_t taxID) const {                  // getParentTaxID, :225-231: the root's children end the walk
    auto it = nodes.find(taxID);
    return (it != nodes.end() && it->second.parentTaxonomyID != 1) ? it->second.parentTaxonomyID : 0;
  }
  const std::string &name(uint32_t taxID) const {          // getScientificName, :233-239
    static const std::string empty;
    auto it = nodes.find(taxID);
    return it != nodes.end() ? it->second.scientificName : empty;
  }
  const std::string &rank(uint32_t taxID) const {          // getRank, :241-247
    static const std::string empty;
    auto it = nodes.find(taxID);
    return it != nodes.end() ? it->second.rank : empty;
  }
  // getLowestCommonAncestor, :185-223. Paths run from the node up to (not including) the root's child boundary, are
  // reversed, ordered by length, and compared position by position over the shortest one. The reference walks parent
  // links without a cycle check; a database with a cycle would hang it, here the walk stops after nodes.size() steps.
  // What the comparison over root-first paths computes is the last element of the longest common prefix of all paths
  // (cut at the shortest one): the prefix shared with the first path can only shrink as further paths are compared.
  uint32_t lca(const uint32_t *taxIDs, uint64_t n) const {
    if (n == 0) return 0;
    constexpr int CAP = 256;
    uint32_t first[CAP], other[CAP];
    const int len0 = path_of(taxIDs[0], first, CAP);
    if (len0 < 0) return lca_deep(taxIDs, n);
    int lcp = len0;
    for (uint64_t k = 1; k < n && lcp > 0; k++) {
      if (taxIDs[k] == taxIDs[k - 1]) continue;            // alignments of one read often share the entry's taxon
      const int len = path_of(taxIDs[k], other, CAP);
      if (len < 0) return lca_deep(taxIDs, n);
      const int m = lcp < len ? lcp : len;
      int i = 0;
      while (i < m && first[len0 - 1 - i] == other[len - 1 - i]) i++;
      lcp = i;
    }
    return lcp > 0 ? first[len0 - lcp] : 0;
  }
  // Same answer for walks longer than 256 levels (or a database whose parent links loop: a walk is cut after nodes.size() + 1
  // ids): the prefix fold of lca() over heap-allocated leaf-first walks.
  void walk_up(uint32_t taxID, std::vector<uint32_t> &ids) const {
    ids.clear();
    for (uint32_t t = taxID; t != 0 && ids.size() <= nodes.size(); t = parent(t)) ids.push_back(t);
  }
  uint32_t lca_deep(const uint32_t *taxIDs, uint64_t n) const {
    std::vector<uint32_t> head, cur;
    walk_up(taxIDs[0], head);
    size_t shared = head.size();                           // how many ids, counted from the root end, every walk so far shares
    for (uint64_t k = 1; k < n && shared > 0; k++) {
      walk_up(taxIDs[k], cur);
      const size_t lim = std::min(shared, cur.size());
      size_t same = 0;
      while (same < lim && head[head.size() - 1 - same] == cur[cur.size() - 1 - same]) same++;
      shared = same;
    }
    return shared ? head[head.size() - shared] : 0;
  }
  // getLineage, :249-265. Walking towards the root the reference puts each node's name in FRONT of the text so far
  // ("; " between, only when there is text already) and empties the text at every node of rank "species"; 131567
  // (cellular organisms) is never named; "." closes a non-empty text once the root is passed. So the text holds the
  // ancestors above the topmost species node (all nodes when there is none): collected leaf-first here, written root-first.
  std::string lineage(uint32_t start) const {
    std::vector<uint32_t> kept;
    bool reached_root = false;
    uint32_t id = start;
    for (size_t steps = 0; steps <= nodes.size() + 1 && !reached_root; steps++) {   // (bounded: a cyclic parent table ends the walk)
      if (id != 131567) {
        kept.push_back(id);
        if (rank(id) == "species") kept.clear();
      }
      id = parent(id);
      reached_root = id == 0;
    }
    // separator after a name <=> some name below it is non-empty (that is when the text was non-empty as it went in front)
    std::vector<uint8_t> text_below(kept.size(), 0);
    bool any = false;
    for (size_t k = 0; k < kept.size(); k++) { text_below[k] = any; any = any || !name(kept[k]).empty(); }
    std::string out;
    for (size_t k = kept.size(); k-- > 0;) {
      out += name(kept[k]);
      if (text_below[k]) out += "; ";
    }
    if (reached_root && !out.empty()) out += '.';
    return out;
  }
};

// ---- IdentifiedTaxonomy (MetagenomicResults.h:32-42), one per read pair, kept compact: the read id lives in an arena
// (the batch's buffers are recycled), genes are references into the database's gene table + a count.
namespace {
struct GeneHit { const kslam_gene *g = nullptr; int count = 1; };
struct PerRead { uint32_t taxonomyID = 0; uint32_t id_len = 0; uint64_t id_off = 0; uint64_t gene_off = 0; uint32_t n_genes = 0; uint32_t has_read = 0; };
}

struct kslam_taxa {
  kslam_sam_db db{};                                       // gene strings are read through it until kslam_taxa_results
  std::vector<PerRead> items;
  std::string ids;
  std::vector<GeneHit> genes;
};

namespace {

struct GeneOps {
  const kslam_sam_db *db;
  std::string_view s(const kslam_gene *g, int w) const { return gene_str(db, g, w); }
  bool equal(const GeneHit &a, const GeneHit &b) const {   // Gene::operator==, GenbankTools.h:82-89
    if (s(a.g, GENE_PROTEIN).size() == 0 && s(b.g, GENE_PROTEIN).size() == 0) return s(a.g, GENE_NAME) == s(b.g, GENE_NAME);
    if (s(a.g, GENE_PROTEIN) == s(b.g, GENE_PROTEIN)) return s(a.g, GENE_PRODUCT) == s(b.g, GENE_PRODUCT);
    return false;
  }
  bool less(const GeneHit &a, const GeneHit &b) const {    // geneSort, GenbankTools.h:116-125
    if (s(a.g, GENE_PROTEIN).size() == 0 && s(b.g, GENE_PROTEIN).size() == 0) return s(a.g, GENE_NAME) < s(b.g, GENE_NAME);
    if (s(a.g, GENE_PROTEIN) == s(b.g, GENE_PROTEIN)) return s(a.g, GENE_PRODUCT) < s(b.g, GENE_PRODUCT);
    return s(a.g, GENE_PROTEIN) < s(b.g, GENE_PROTEIN);
  }
};

void xml_escape(std::string &out, std::string_view in) {   // correctXML, MetagenomicResults.h:276-301
  for (char ch : in) switch (ch) {
    case '<': out += "&lt;"; break;
    case '>': out += "&gt;"; break;
    case '&': out += "&amp;"; break;
    case '\'': out += "&apos;"; break;
    case '"': out += "&quot;"; break;
    default: out.push_back(ch);
  }
}

struct Combined { uint32_t taxonomyID = 0; std::vector<std::string_view> reads; std::vector<GeneHit> genes; };

}  // namespace

// getResultFromPairedOverlaps (MetagenomicResults.h:88-111) for every per-read record of the batch, in order
int kslam_host::taxa_add_batch(kslam_taxa *taxa, const kslam_taxdb *taxdb, const Ctx &c, const std::vector<ReadPair> &rp, uint32_t threads) {
  try {
    if (taxa->items.empty()) taxa->db = *c.db;
    else if (taxa->db.genes != c.db->genes || taxa->db.gene_strings != c.db->gene_strings) return KSLAM_ERR_STATE;   // one database per run
    const GeneOps ops{c.db};
    struct Part { std::vector<PerRead> items; std::string ids; std::vector<GeneHit> genes; };
    std::vector<Part> parts(std::max(1u, threads));
    parallel_ranges(threads, rp.size(), [&](uint32_t t, size_t lo, size_t hi) {
      Part &part = parts[t];
      std::vector<uint32_t> taxIDs;
      std::vector<GeneHit> genes;
      for (size_t i = lo; i < hi; i++) {
        const ReadPair &read = rp[i];
        PerRead r;
        if (read.pairs.size()) {
          taxIDs.clear(); genes.clear();
          for (const POv &a : read.pairs) {
            taxIDs.push_back(c.db->taxonomy_ids ? c.db->taxonomy_ids[a.entry] : 0);
            if (const kslam_gene *g = best_gene(c.db, a.entry, a.refStart, a.refEnd)) genes.push_back(GeneHit{g, 1});
          }
          std::sort(genes.begin(), genes.end(), [&](const GeneHit &x, const GeneHit &y) { return ops.less(x, y); });
          genes.erase(std::unique(genes.begin(), genes.end(), [&](const GeneHit &x, const GeneHit &y) { return ops.equal(x, y); }), genes.end());
          r.has_read = 1;
          r.id_off = part.ids.size();
          r.id_len = (uint32_t)(c.reads->id_offs[read.r1Pos + 1] - c.reads->id_offs[read.r1Pos]);
          part.ids.append(c.reads->ids + c.reads->id_offs[read.r1Pos], r.id_len);
          r.gene_off = part.genes.size(); r.n_genes = (uint32_t)genes.size();
          part.genes.insert(part.genes.end(), genes.begin(), genes.end());
          r.taxonomyID = taxdb->lca(taxIDs.data(), taxIDs.size());
        }
        part.items.push_back(r);
      }
    });
    for (Part &part : parts) {
      const uint64_t id_base = taxa->ids.size(), gene_base = taxa->genes.size();
      taxa->ids += part.ids;
      taxa->genes.insert(taxa->genes.end(), part.genes.begin(), part.genes.end());
      for (PerRead r : part.items) { r.id_off += id_base; r.gene_off += gene_base; taxa->items.push_back(r); }
    }
    return KSLAM_OK;
  } catch (const std::exception &) { return KSLAM_ERR_NOMEM; }
}

extern "C" {

int kslam_taxdb_open(const char *path, kslam_taxdb **out) {          // TaxonomyDB(inFileName), TaxonomyDatabase.h:87-93
  if (!path || !out) return KSLAM_ERR_ARG;
  kslam_taxdb *db = nullptr;
  try {
    db = new kslam_taxdb();
    db->read_index(path);
    db->finalize();
    *out = db;
    return KSLAM_OK;
  } catch (const std::exception &) { delete db; return KSLAM_ERR_ARG; }   // unreadable file or a line stoi rejects
}

int kslam_taxdb_build(const char *names_dmp, const char *nodes_dmp, const char *out_path) {   // writeTaxonomyIndex, :153-164
  if (!names_dmp || !nodes_dmp || !out_path) return KSLAM_ERR_ARG;
  try {
    kslam_taxdb db;
    db.parse_nodes(nodes_dmp);
    db.parse_names(names_dmp);
    std::ofstream out(out_path);
    if (!out.is_open()) return KSLAM_ERR_ARG;
    for (auto &e : db.nodes) out << e.first << "\n" << e.second.parentTaxonomyID << "\n" << e.second.scientificName << "\n" << e.second.rank << "\n";
    out.close();
    return out.fail() ? KSLAM_ERR_STATE : KSLAM_OK;
  } catch (const std::exception &) { return KSLAM_ERR_ARG; }
}

uint64_t kslam_taxdb_size(const kslam_taxdb *db) { return db ? db->nodes.size() : 0; }
uint32_t kslam_taxdb_lca(const kslam_taxdb *db, const uint32_t *tax_ids, uint64_t n) { return (db && (tax_ids || !n)) ? db->lca(tax_ids, n) : 0; }

int kslam_taxdb_lineage(const kslam_taxdb *db, uint32_t tax_id, char **text, uint64_t *len) {
  if (!db || !text) return KSLAM_ERR_ARG;
  const std::string s = db->lineage(tax_id);
  *text = dup_text(s);
  if (len) *len = s.size();
  return *text ? KSLAM_OK : KSLAM_ERR_NOMEM;
}
int kslam_taxdb_name(const kslam_taxdb *db, uint32_t tax_id, char **text, uint64_t *len) {
  if (!db || !text) return KSLAM_ERR_ARG;
  const std::string &s = db->name(tax_id);
  *text = dup_text(s);
  if (len) *len = s.size();
  return *text ? KSLAM_OK : KSLAM_ERR_NOMEM;
}
void kslam_taxdb_close(kslam_taxdb *db) { delete db; }

int kslam_taxa_create(kslam_taxa **out) {
  if (!out) return KSLAM_ERR_ARG;
  *out = new (std::nothrow) kslam_taxa();
  return *out ? KSLAM_OK : KSLAM_ERR_NOMEM;
}
void kslam_taxa_destroy(kslam_taxa *taxa) { delete taxa; }

// End of the run, SLAM.h:256-265: <out>_PerRead from the per-read results in batch order; then one record per taxon.
// The reference does all of this on one thread; for runs of 10^8 read pairs that would take longer than the alignment.
// Here only combineTaxonomies' sort stays sequential — its permutation decides which of two genes that compare equal
// represents the group, so it has to be THE std::sort; it runs over (taxon, index) pairs of 8 bytes instead of the
// records (the permutation of std::sort depends on the comparisons only). Everything else is per taxon or per read range
// and runs on all host threads: the _PerRead text, the per-taxon merges, the sorts of the read names (parallel chunks +
// merges; names that compare equal are identical, so any algorithm gives the reference's sequence) and the XML text.
int kslam_taxa_results(kslam_taxa *taxa, const kslam_taxdb *taxdb, uint32_t num_reads, char **per_read, uint64_t *per_read_len,
                       char **xml, uint64_t *xml_len, char **abbreviated, uint64_t *abbreviated_len) {
  if (!taxa || !taxdb) return KSLAM_ERR_ARG;
  try {
    const GeneOps ops{&taxa->db};
    uint32_t threads = std::max(1u, std::thread::hardware_concurrency());
    if (threads > 64) threads = 64;
    if (const char *e = getenv("KSLAM_HOST_THREADS")) threads = (uint32_t)std::max(1, atoi(e));
    size_t par_min = 1u << 16;                             // smaller inputs are handled on one thread (KSLAM_HOST_PAR_MIN: test hook)
    if (const char *e = getenv("KSLAM_HOST_PAR_MIN")) par_min = (size_t)strtoull(e, nullptr, 10);
    const std::vector<PerRead> &items = taxa->items;
    auto id_of = [&](const PerRead &r) { return std::string_view(taxa->ids.data() + r.id_off, r.id_len); };
    auto join = [](std::vector<std::string> &parts, char **text, uint64_t *len) {
      size_t total = 0;
      for (auto &p : parts) total += p.size();
      char *buf = (char *)malloc(total + 1);
      if (!buf) return false;
      size_t at = 0;
      for (auto &p : parts) { memcpy(buf + at, p.data(), p.size()); at += p.size(); }
      buf[total] = 0;
      *text = buf;
      if (len) *len = total;
      return true;
    };
    if (per_read) {                                        // writePerReadResults, MetagenomicResults.h:455-463
      std::vector<std::string> parts(threads);
      parallel_ranges(threads, items.size(), [&](uint32_t t, size_t lo, size_t hi) {
        std::string &o = parts[t];
        for (size_t i = lo; i < hi; i++)
          if (items[i].has_read) { o += id_of(items[i]); o.push_back('\t'); o += std::to_string(items[i].taxonomyID); o.push_back('\n'); }
      });
      if (!join(parts, per_read, per_read_len)) return KSLAM_ERR_NOMEM;
    }
    if (!xml && !abbreviated) return KSLAM_OK;
    // combineTaxonomies, :149-176. The loop skips the first element and starts with testTaxID = 0: reads without a taxon
    // (id 0) sort first and are dropped, and when there are none the very first record of the sorted vector is left out
    // of its range (or its taxon dropped, if it was alone) — kept as is.
    struct Key { uint32_t taxonomyID, index; };
    std::vector<Key> sorted(items.size());
    parallel_ranges(threads, items.size(), [&](uint32_t, size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; i++) sorted[i] = Key{items[i].taxonomyID, (uint32_t)i};
    });
    if (items.size() > UINT32_MAX) return KSLAM_ERR_ARG;
    std::sort(sorted.begin(), sorted.end(), [](const Key &i, const Key &j) { return i.taxonomyID < j.taxonomyID; });
    std::vector<std::pair<size_t, size_t>> ranges;          // [begin, end) of every taxon that is kept
    if (!sorted.empty()) {
      uint32_t testTaxID = 0;
      size_t start = 0;
      for (size_t tax = 1; tax < sorted.size(); tax++)
        if (sorted[tax].taxonomyID != testTaxID) {
          if (testTaxID != 0) ranges.push_back({start, tax});
          testTaxID = sorted[tax].taxonomyID;
          start = tax;
        }
      if (sorted[start].taxonomyID != 0) ranges.push_back({start, sorted.size()});
    }
    std::vector<Combined> results(ranges.size());
    // sorted sequence of a big vector on all threads: sorted chunks, then rounds of pairwise merges
    auto sort_views = [&](std::vector<std::string_view> &v) {
      const size_t n = v.size();
      if (threads <= 1 || n < par_min || n < 4 * (size_t)threads) { std::sort(v.begin(), v.end()); return; }
      uint32_t chunks = 1;
      while (chunks * 2 <= threads) chunks *= 2;
      auto cut = [&](uint32_t c) { return n * c / chunks; };
      parallel_threads(chunks, [&](uint32_t c) { std::sort(v.begin() + cut(c), v.begin() + cut(c + 1)); });
      for (uint32_t width = 1; width < chunks; width *= 2)
        parallel_threads(chunks / (2 * width), [&](uint32_t m) {
          const uint32_t c = m * 2 * width;
          std::inplace_merge(v.begin() + cut(c), v.begin() + cut(c + width), v.begin() + cut(c + 2 * width));
        });
    };
    auto combine = [&](size_t which, bool inner_parallel) { // combineRangeOfIdentifiedTaxonomy, :117-143
      Combined &t = results[which];
      const size_t begin = ranges[which].first, end = ranges[which].second;
      t.taxonomyID = sorted[begin].taxonomyID;
      size_t n_genes = 0;
      for (size_t k = begin; k < end; k++) n_genes += items[sorted[k].index].n_genes;
      t.genes.reserve(n_genes); t.reads.reserve(end - begin);
      for (size_t k = begin; k < end; k++) {
        const PerRead &r = items[sorted[k].index];
        t.genes.insert(t.genes.end(), taxa->genes.begin() + r.gene_off, taxa->genes.begin() + r.gene_off + r.n_genes);
        if (r.has_read) t.reads.push_back(id_of(r));
      }
      std::sort(t.genes.begin(), t.genes.end(), [&](const GeneHit &x, const GeneHit &y) { return ops.less(x, y); });
      size_t n_runs = 0;                                      // equal neighbours collapse into the first of their run, which counts them
      for (size_t k = 0; k < t.genes.size(); k++) {
        if (n_runs && ops.equal(t.genes[n_runs - 1], t.genes[k])) t.genes[n_runs - 1].count++;
        else { if (n_runs != k) t.genes[n_runs] = t.genes[k]; n_runs++; }
      }
      t.genes.resize(n_runs);
      // sortResults, :254-275, the per-entry part (the order of the entries is decided below)
      if (inner_parallel) sort_views(t.reads); else std::sort(t.reads.begin(), t.reads.end());
      auto by_count = [&](const GeneHit &i, const GeneHit &j) {
        if (i.count == j.count) {
          if (i.g->cds_start == j.g->cds_start) return ops.s(i.g, GENE_LOCUS) < ops.s(j.g, GENE_LOCUS);
          return i.g->cds_start < j.g->cds_start;
        }
        return i.count > j.count;
      };
      std::sort(t.genes.begin(), t.genes.end(), by_count);  // (writeAbbreviatedResultsFile sorts once more AFTER the XML is out: no visible effect)
    };
    // taxa with very many reads one at a time with all threads inside; the rest handed out one taxon per thread
    const size_t big = std::max<size_t>(4 * par_min, getenv("KSLAM_HOST_PAR_MIN") ? 0 : sorted.size() / 8);
    std::vector<size_t> small;
    for (size_t i = 0; i < ranges.size(); i++) {
      if (ranges[i].second - ranges[i].first >= big) combine(i, true);
      else small.push_back(i);
    }
    {
      std::atomic<size_t> next{0};
      parallel_threads(small.size() < 2 ? 1u : threads, [&](uint32_t) { for (size_t k = next++; k < small.size(); k = next++) combine(small[k], false); });
    }
    auto by_size = [](const Combined &i, const Combined &j) {
      if (i.reads.size() == j.reads.size()) return i.taxonomyID < j.taxonomyID;
      return i.reads.size() > j.reads.size();
    };
    std::sort(results.begin(), results.end(), by_size);     // a total order (ids are unique): sorting twice changes nothing
    if (xml) {                                             // getXML, :302-369
      std::vector<std::string> parts(results.size());
      auto reads_block = [&](const Combined &entry, size_t lo, size_t hi, std::string &o) {
        for (size_t i = lo; i < hi; i++) { o += "    <read>"; xml_escape(o, entry.reads[i]); o += "</read>\n"; }
      };
      auto entry_xml = [&](size_t which, bool inner_parallel) {
        const Combined &entry = results[which];
        std::string &o = parts[which];
        o += "<taxon>\n  <abundance numReads=\"";
        o += std::to_string(entry.reads.size());
        o += "\">";
        o += std::to_string(entry.reads.size() * 100.0 / num_reads);
        o += "</abundance>\n  <taxonomyID>";
        o += std::to_string(entry.taxonomyID);
        o += "</taxonomyID>\n  <lineage>";
        xml_escape(o, taxdb->lineage(entry.taxonomyID));
        o += "</lineage>\n  <name>";
        xml_escape(o, taxdb->name(entry.taxonomyID));
        o += "</name>\n  <genes>\n";
        for (const GeneHit &gene : entry.genes) {
          o += "    <gene protein=\""; xml_escape(o, ops.s(gene.g, GENE_PROTEIN));
          o += "\" locus=\""; xml_escape(o, ops.s(gene.g, GENE_LOCUS));
          o += "\" product=\""; xml_escape(o, ops.s(gene.g, GENE_PRODUCT));
          o += "\" GeneID=\""; o += std::to_string(gene.g->gene_id);
          o += "\" reference=\""; xml_escape(o, ops.s(gene.g, GENE_REFERENCE));
          o += "\" numReads=\""; o += std::to_string(gene.count);
          o += "\" cdsStart=\""; o += std::to_string(gene.g->cds_start);
          o += "\" cdsEnd=\""; o += std::to_string(gene.g->cds_stop);
          o += "\">"; xml_escape(o, ops.s(gene.g, GENE_NAME));
          o += "</gene>\n";
        }
        o += "  </genes>\n  <reads>\n";
        if (inner_parallel) {
          std::vector<std::string> sub(threads);
          parallel_ranges(threads, entry.reads.size(), [&](uint32_t t, size_t lo, size_t hi) { reads_block(entry, lo, hi, sub[t]); });
          for (auto &x : sub) o += x;
        } else reads_block(entry, 0, entry.reads.size(), o);
        o += "  </reads>\n</taxon>\n";
      };
      std::vector<size_t> small_x;
      for (size_t i = 0; i < results.size(); i++) {
        if (results[i].reads.size() >= big) entry_xml(i, true);
        else small_x.push_back(i);
      }
      std::atomic<size_t> next{0};
      parallel_threads(small_x.size() < 2 ? 1u : threads, [&](uint32_t) { for (size_t k = next++; k < small_x.size(); k = next++) entry_xml(small_x[k], false); });
      if (!join(parts, xml, xml_len)) return KSLAM_ERR_NOMEM;
    }
    if (abbreviated) {                                     // writeAbbreviatedResultsFile, :237-249 (ostream << double = %g)
      std::string o;
      char num[64];
      for (const Combined &entry : results) {
        o += taxdb->name(entry.taxonomyID);
        o.push_back('\t');
        snprintf(num, sizeof num, "%g", entry.reads.size() * 100.0 / num_reads);
        o += num;
        o.push_back('\n');
      }
      *abbreviated = dup_text(o);
      if (abbreviated_len) *abbreviated_len = o.size();
      if (!*abbreviated) return KSLAM_ERR_NOMEM;
    }
    return KSLAM_OK;
  } catch (const std::exception &) { return KSLAM_ERR_NOMEM; }
}

}  // extern "C"
