// genbank.cu — the database side of the drop-in (host code): building a GenbankIndex from GenBank flat files or FASTA
// files and reading / writing DIR/database, the file `--db` points at.
//   createIndexFromGBFF + parseSection      /root/reference/src/GenbankTools.h:348-527   (--parse-genbank)
//   createIndexFromFASTA                    GenbankTools.h:224-260                       (--parse-fasta)
//   writeIndexToBoostSerial / getIndexFromBoostSerial   GenbankTools.h:201-205,336-344   (Boost text archive)
// The parsers follow the reference field by field, including what it does by accident (the taxon id is found because
// stoul stops at the closing quote after an unsigned wrap-around of the substring length; a qualifier found anywhere in
// a feature's text wins; every line of the ORIGIN block is a "section" of its own whose tag is the base counter).
// The archive grammar is the one SURVEY.md App. B.1 spells out. This image has no Boost headers (the reference's writer cannot
// be compiled here), but the REAL Boost.Serialization 1.78 library is present as a header-less .so: oracle/boost_archive_probe.cpp
// drives it, oracle/ref_shim_boost runs the reference's own writeIndexToBoostSerial / getIndexFromBoostSerial on it, and
// tests/test_database_format.py checks this file's writer against both byte for byte, this file's reader on their archives, and the
// reference's reader on this file's archives. The parsers are pinned against the reference's own createIndexFromGBFF / createIndexFromFASTA
// (tests/test_taxon_host.py, through oracle/_ref).
#include "common.cuh"
#include <unistd.h>
#include <sys/stat.h>
#include <sys/mman.h>
#include <fcntl.h>
#include "host_stages.h"
#include <algorithm>
#include <atomic>
#include <climits>
#include <ctype.h>
#include <stdexcept>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <thread>

namespace {

struct Gene {                                              // Gene + CDS, GenbankTools.h:47-110
  std::string geneName, locusTag, proteinID, product, referenceSequence;
  uint32_t geneID = 0, start = 0, stop = 0;
  bool complement = false;
};
struct Entry {                                             // GenbankEntry, GenbankTools.h:136-164 (serialised members + definition)
  std::string bases, locusTag, definition;
  uint32_t taxonomyID = 0, genbankID = 0;
  bool isPlasmid = false, is16S = false;
  std::vector<Gene> genes;
};

thread_local std::string g_error;

bool read_file(const char *path, std::string &data) {
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  if (fseek(f, 0, SEEK_END) == 0) {                        // regular file: one allocation, one read
    const long size = ftell(f);
    rewind(f);
    if (size > 0) {
      data.resize((size_t)size);
      const size_t got = fread(&data[0], 1, (size_t)size, f);
      data.resize(got);
    }
  }
  char buf[1 << 16];                                       // whatever is left (pipes, files that grew)
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
  const bool ok = !ferror(f);
  fclose(f);
  return ok;
}

typedef std::string::const_iterator It;
inline It skip_to(It from, It end, bool (*pred)(char)) { return std::find_if(from, end, pred); }
bool is_space(char c) { return c == ' '; }
bool not_space(char c) { return c != ' '; }
bool is_digit(char c) { return isdigit((unsigned char)c) != 0; }
bool not_digit(char c) { return isdigit((unsigned char)c) == 0; }

// text between `key` and the next double quote, if both exist (the pattern of GenbankTools.h:415-468)
bool qualifier(const std::string &field, size_t startPos, size_t keyLen, std::string &out) {
  if (startPos == std::string::npos) return false;
  startPos += keyLen;
  const size_t endPos = field.find('"', startPos);
  if (endPos == std::string::npos || startPos >= field.size()) return false;
  out = field.substr(startPos, endPos - startPos);
  return true;
}

void parse_section(const std::string &field, Entry &entry) {          // parseSection, GenbankTools.h:348-476
  It start = skip_to(field.begin(), field.end(), not_space);
  if (start == field.end()) return;
  It stop = skip_to(start, field.end(), is_space);
  const std::string tag(start, stop);
  start = skip_to(stop, field.end(), not_space);
  if (tag == "VERSION") {
    stop = skip_to(start, field.end(), is_space);
    entry.locusTag = std::string(start, stop);
    start = skip_to(stop, field.end(), is_digit);
    try { entry.genbankID = (uint32_t)stoul(std::string(start, field.end())); } catch (...) {}
  } else if (tag == "DEFINITION") {
    entry.definition = std::string(start, field.end());
  } else if (tag == "source") {
    size_t startPos = field.find("/db_xref=\"taxon:");
    const size_t endPos = field.find('"', startPos);       // the OPENING quote: endPos < startPos + 16, the length below wraps
    if (startPos != std::string::npos && endPos != std::string::npos) {
      startPos += 16;
      if (startPos < field.size()) try { entry.taxonomyID = (uint32_t)stoul(field.substr(startPos, endPos - startPos)); } catch (...) {}
    }
  } else if (tag == "CDS" || tag == "tRNA" || tag == "gene") {
    Gene gene;
    start = skip_to(start, field.end(), is_digit);
    stop = skip_to(start, field.end(), not_digit);
    try { gene.start = (uint32_t)stoul(std::string(start, stop)); } catch (...) {}
    start = skip_to(stop, field.end(), is_digit);
    stop = skip_to(start, field.end(), not_digit);
    try { gene.stop = (uint32_t)stoul(std::string(start, stop)); } catch (...) {}
    qualifier(field, field.find("/product=\""), 10, gene.product);
    qualifier(field, field.rfind("/protein_id=\""), 13, gene.proteinID);
    qualifier(field, field.find("/locus_tag=\""), 12, gene.locusTag);
    std::string id;
    if (qualifier(field, field.find("GeneID:"), 7, id)) gene.geneID = (uint32_t)std::stoul(id);   // not guarded in the reference either: throws
    qualifier(field, field.find("/gene=\""), 7, gene.geneName);
    gene.referenceSequence = entry.locusTag;
    entry.genes.push_back(std::move(gene));
  } else if (tag.size() && is_digit(tag[0])) {
    for (; start != field.end(); ++start)
      if (*start != ' ') entry.bases.push_back((char)toupper((unsigned char)*start));
  }
}

}  // namespace

// GenbankIndex held flat: what kslam_load_genomes, kslam_sam_db and the archive writer take without another copy.
struct kslam_index {
  std::string bases; std::vector<uint64_t> offs{0};
  // bases of an index that came from the side-car cache live in the mapping of that file, not in `bases`
  const char *mapped_bases = nullptr; void *map = nullptr; size_t map_len = 0;
  const char *bases_data() const { return mapped_bases ? mapped_bases : bases.data(); }
  std::string locus; std::vector<uint64_t> locus_offs{0};
  std::vector<uint32_t> taxonomy_ids, genbank_ids;
  std::vector<uint8_t> is_plasmid, is_16s;
  std::vector<kslam_gene> genes; std::vector<uint64_t> gene_offs{0}; std::string gene_strings;

  void add(const Entry &e) { add(e, e.bases); }
  void add(const Entry &e, std::string_view entry_bases) {   // the bases may live elsewhere (the archive reader passes a view)
    bases.append(entry_bases.data(), entry_bases.size()); offs.push_back(bases.size());
    locus += e.locusTag; locus_offs.push_back(locus.size());
    taxonomy_ids.push_back(e.taxonomyID); genbank_ids.push_back(e.genbankID);
    is_plasmid.push_back(e.isPlasmid); is_16s.push_back(e.is16S);
    for (const Gene &g : e.genes) {
      kslam_gene k;
      k.cds_start = g.start; k.cds_stop = g.stop; k.gene_id = g.geneID; k.complement = g.complement;
      const std::string *s[5] = {&g.geneName, &g.locusTag, &g.proteinID, &g.product, &g.referenceSequence};
      for (int i = 0; i < 5; i++) { k.str_offs[i] = gene_strings.size(); gene_strings += *s[i]; }
      k.str_offs[5] = gene_strings.size();
      genes.push_back(k);
    }
    gene_offs.push_back(genes.size());
  }
  uint64_t n() const { return offs.size() - 1; }
  mutable kslam_gene_index *gene_index = nullptr;          // built on the first kslam_index_db
  ~kslam_index() { delete gene_index; if (map) munmap(map, map_len); }
};

namespace {

// ---- Boost text archive: space-separated tokens after "22 serialization::archive <libver>"; the FIRST
// object of every class type is preceded by "<tracking> <version>" = "0 0"; a vector is "<count> <item_version>" then the
// items; a string is "<len>", ONE space, then len raw bytes; bool is 0 / 1. Serialised members, in order:
//   GenbankIndex{entries}  GenbankEntry{bases taxonomyID genbankID isPlasmid is16S locusTag genes}
//   Gene{geneName locusTag proteinID product referenceSequence geneID codingSequence}  CDS{start stop complement}
const char kArchiveHeader[] = "22 serialization::archive";

struct ArchiveReader {
  const std::string &d; size_t p; unsigned seen = 0;
  uint64_t number() {
    while (p < d.size() && (d[p] == ' ' || d[p] == '\n' || d[p] == '\r' || d[p] == '\t')) p++;
    size_t q = p;
    uint64_t v = 0;
    while (q < d.size() && is_digit(d[q])) { v = v * 10 + (uint64_t)(d[q] - '0'); q++; }
    if (q == p) throw std::runtime_error("expected a number at byte " + std::to_string(p));
    p = q;
    return v;
  }
  std::string_view view() {
    const uint64_t n = number();
    p += 1;                                                // exactly one separator, then n raw bytes (they may contain spaces)
    if (p > d.size() || n > d.size() - p) throw std::runtime_error("string runs past the end of the archive");
    const std::string_view v(d.data() + p, n);
    p += n;
    return v;
  }
  void string(std::string &out) { const std::string_view v = view(); out.assign(v.data(), v.size()); }
  void class_info(unsigned bit) { if (!(seen & bit)) { seen |= bit; number(); number(); } }
  uint64_t vector(unsigned bit) { class_info(bit); const uint64_t count = number(); number(); return count; }
};

// ---- side-car cache of DIR/database -------------------------------------------------------------------
// DIR/database is a Boost TEXT archive (GenbankTools.h:197-205,336-344): reading it means tokenising every number and
// copying every base. `<database>.kslam` holds the same GenbankIndex as flat binary arrays, keyed by the size and
// modification time of the archive it was made from; it is mapped, the small tables are copied and the bases are used in
// place. It is written (best effort, silently skipped in a read-only directory or with KSLAM_NO_INDEX_CACHE set) after
// the first successful parse and ignored whenever its key no longer matches. Only the HOST side is cached: the device side
// (2-bit planes, sorted genome k-mer list, prefilter bitmap) is ~2 bytes per base and the B200 rebuilds it from the bases at
// ~0.15 s per Gbp — faster than those bytes come in over PCIe from pageable memory.
struct CacheHeader {
  char magic[8];
  uint64_t src_size, src_mtime_ns, n_entries, n_bases, n_locus, n_genes, n_gene_strings;
};
const char kCacheMagic[8] = {'K', 'S', 'L', 'A', 'M', 'I', 'X', '1'};

bool source_key(const char *path, uint64_t *size, uint64_t *mtime_ns) {
  struct stat st;
  if (stat(path, &st) != 0 || !S_ISREG(st.st_mode)) return false;
  *size = (uint64_t)st.st_size; *mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
  return true;
}
size_t pad8(size_t n) { return (n + 7) & ~(size_t)7; }

kslam_index *cache_load(const std::string &cache_path, uint64_t src_size, uint64_t src_mtime_ns) {
  const int fd = open(cache_path.c_str(), O_RDONLY);
  if (fd < 0) return nullptr;
  struct stat st;
  if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(CacheHeader)) { close(fd); return nullptr; }
  void *map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return nullptr;
  const CacheHeader *h = (const CacheHeader *)map;
  const size_t n = (size_t)h->n_entries;
  size_t need = sizeof(CacheHeader) + 3 * pad8((n + 1) * 8) + 2 * pad8(n * 4) + 2 * pad8(n) + pad8((size_t)h->n_genes * sizeof(kslam_gene)) +
                pad8((size_t)h->n_locus) + pad8((size_t)h->n_gene_strings) + (size_t)h->n_bases;
  if (memcmp(h->magic, kCacheMagic, 8) != 0 || h->src_size != src_size || h->src_mtime_ns != src_mtime_ns || need != (size_t)st.st_size) {
    munmap(map, (size_t)st.st_size);
    return nullptr;
  }
  kslam_index *ix = new kslam_index();
  ix->map = map; ix->map_len = (size_t)st.st_size;
  const char *p = (const char *)map + sizeof(CacheHeader);
  auto take = [&](void *dst, size_t bytes) { memcpy(dst, p, bytes); p += pad8(bytes); };
  ix->offs.resize(n + 1); take(ix->offs.data(), (n + 1) * 8);
  ix->locus_offs.resize(n + 1); take(ix->locus_offs.data(), (n + 1) * 8);
  ix->gene_offs.resize(n + 1); take(ix->gene_offs.data(), (n + 1) * 8);
  ix->taxonomy_ids.resize(n); take(ix->taxonomy_ids.data(), n * 4);
  ix->genbank_ids.resize(n); take(ix->genbank_ids.data(), n * 4);
  ix->is_plasmid.resize(n); take(ix->is_plasmid.data(), n);
  ix->is_16s.resize(n); take(ix->is_16s.data(), n);
  ix->genes.resize((size_t)h->n_genes); take(ix->genes.data(), (size_t)h->n_genes * sizeof(kslam_gene));
  ix->locus.resize((size_t)h->n_locus); take(&ix->locus[0], (size_t)h->n_locus);
  ix->gene_strings.resize((size_t)h->n_gene_strings); take(&ix->gene_strings[0], (size_t)h->n_gene_strings);
  ix->mapped_bases = p;
  // the key matched, but the tables are still a file's contents: every offset table must start at 0, never decrease and
  // end at the size of what it indexes, or the side-car is dropped like a stale one (the archive is parsed instead)
  auto table_ok = [n](const std::vector<uint64_t> &t, uint64_t total) {
    if (t[0] != 0 || t[n] != total) return false;
    for (size_t i = 0; i < n; i++) if (t[i] > t[i + 1]) return false;
    return true;
  };
  bool ok = table_ok(ix->offs, h->n_bases) && table_ok(ix->locus_offs, h->n_locus) && table_ok(ix->gene_offs, h->n_genes);
  for (size_t g = 0; ok && g < ix->genes.size(); g++) {
    const uint64_t *so = ix->genes[g].str_offs;
    for (int k = 0; k < 5; k++) ok = ok && so[k] <= so[k + 1];
    ok = ok && so[5] <= h->n_gene_strings;
  }
  if (!ok) { delete ix; return nullptr; }
  return ix;
}

void cache_store(const kslam_index *ix, const std::string &cache_path, uint64_t src_size, uint64_t src_mtime_ns) {
  const std::string tmp = cache_path + ".tmp" + std::to_string((long)getpid());
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  CacheHeader h;
  memcpy(h.magic, kCacheMagic, 8);
  const size_t n = ix->n();
  h.src_size = src_size; h.src_mtime_ns = src_mtime_ns; h.n_entries = n; h.n_bases = ix->offs[n]; h.n_locus = ix->locus.size();
  h.n_genes = ix->genes.size(); h.n_gene_strings = ix->gene_strings.size();
  bool ok = fwrite(&h, sizeof h, 1, f) == 1;
  static const char zeros[8] = {0};
  auto put = [&](const void *src, size_t bytes) {
    if (bytes) ok = ok && fwrite(src, 1, bytes, f) == bytes;
    if (pad8(bytes) != bytes) ok = ok && fwrite(zeros, 1, pad8(bytes) - bytes, f) == pad8(bytes) - bytes;
  };
  put(ix->offs.data(), (n + 1) * 8); put(ix->locus_offs.data(), (n + 1) * 8); put(ix->gene_offs.data(), (n + 1) * 8);
  put(ix->taxonomy_ids.data(), n * 4); put(ix->genbank_ids.data(), n * 4); put(ix->is_plasmid.data(), n); put(ix->is_16s.data(), n);
  put(ix->genes.data(), ix->genes.size() * sizeof(kslam_gene)); put(ix->locus.data(), ix->locus.size());
  put(ix->gene_strings.data(), ix->gene_strings.size());
  if (h.n_bases) ok = ok && fwrite(ix->bases_data(), 1, (size_t)h.n_bases, f) == (size_t)h.n_bases;
  ok = fclose(f) == 0 && ok;
  if (!ok || rename(tmp.c_str(), cache_path.c_str()) != 0) unlink(tmp.c_str());
}

}  // namespace

extern "C" {

const char *kslam_index_error(void) { return g_error.c_str(); }
void kslam_index_free(kslam_index *index) { delete index; }

// One record: the lines of data[from, to), which end with the "//" line (or with the end of the file, for whatever follows
// the last "//": parsed like the reference does, then dropped). Returns true when the "//" line closed the entry.
static bool parse_genbank_record(const std::string &data, size_t from, size_t to, Entry &entry) {
  std::string line, section;
  for (size_t pos = from; pos < to;) {                     // std::getline: '\n' ends a line, a '\r' stays in it
    size_t nl = data.find('\n', pos);
    if (nl == std::string::npos || nl > to) nl = to;
    line.assign(data, pos, nl - pos);
    pos = nl + 1;
    if (line.size() == 0) continue;
    const size_t startPos = line.find_first_not_of(' ');
    if (startPos < 12) {
      parse_section(section, entry);
      section = line;
      if (line == "//") {
        std::sort(entry.genes.begin(), entry.genes.end(), [](const Gene &i, const Gene &j) {
          if (i.start == j.start) return i.proteinID.size() > j.proteinID.size();
          return i.start < j.start;
        });
        auto it = std::unique(entry.genes.begin(), entry.genes.end(), [](const Gene &i, const Gene &j) { return i.start == j.start; });
        entry.genes.resize(std::distance(entry.genes.begin(), it));
        return true;
      }
    } else if (startPos == std::string::npos) continue;
    else section.append(line.substr(startPos - 1));        // continuation line: one space + its text
  }
  return false;
}

// createIndexFromGBFF, :481-527. The reference walks every file line by line with one entry and one section in flight; a
// line that is exactly "//" pushes the entry and starts a fresh one, and the section it leaves behind ("//") parses to
// nothing — so records are independent and are parsed on all host threads here, files taken in groups of at most ~1 GB,
// entries appended in file order.
int kslam_index_parse_genbank(const char *const *paths, uint64_t n_paths, kslam_index **out) {
  if (!paths || !out) return KSLAM_ERR_ARG;
  kslam_index *index = nullptr;
  try {
    index = new kslam_index();
    const uint32_t threads = std::min(64u, std::max(1u, std::thread::hardware_concurrency()));
    for (uint64_t f0 = 0; f0 < n_paths;) {
      std::vector<std::string> datas;
      size_t group_bytes = 0;
      uint64_t f1 = f0;
      while (f1 < n_paths && (f1 == f0 || group_bytes < (1ull << 30))) {
        datas.emplace_back();
        if (!read_file(paths[f1], datas.back())) throw std::runtime_error(std::string("unable to open index file ") + paths[f1]);
        group_bytes += datas.back().size();
        f1++;
      }
      struct Chunk { uint32_t file; size_t from, to; Entry entry; bool closed = false; std::string error; };
      std::vector<Chunk> chunks;
      for (uint32_t k = 0; k < datas.size(); k++) {
        const std::string &d = datas[k];
        size_t from = 0;
        for (size_t p = d.find("//"); p != std::string::npos; p = d.find("//", p + 2)) {
          const bool line_start = p == 0 || d[p - 1] == '\n', line_end = p + 2 == d.size() || d[p + 2] == '\n';
          if (!line_start || !line_end) continue;
          const size_t to = std::min(p + 3, d.size());
          if (p < from) continue;                          // "///..." overlap guard
          chunks.push_back(Chunk{k, from, to});
          from = to;
        }
        if (from < d.size()) chunks.push_back(Chunk{k, from, d.size()});   // after the last "//": parsed, never pushed
      }
      std::atomic<size_t> next{0};
      kslam_host::parallel_threads(chunks.size() < 2 ? 1u : threads, [&](uint32_t) {
        for (size_t c = next++; c < chunks.size(); c = next++) {
          Chunk &ch = chunks[c];
          try { ch.closed = parse_genbank_record(datas[ch.file], ch.from, ch.to, ch.entry); }
          catch (const std::exception &e) { ch.error = e.what(); if (ch.error.empty()) ch.error = "parse error"; }
        }
      });
      for (Chunk &ch : chunks) {
        if (!ch.error.empty()) throw std::runtime_error(ch.error + " in " + paths[f0 + ch.file]);
        if (ch.closed) index->add(ch.entry);
        ch.entry = Entry();
      }
      f0 = f1;
    }
    *out = index;
    return KSLAM_OK;
  } catch (const std::exception &e) { g_error = e.what(); delete index; return KSLAM_ERR_ARG; }
}

int kslam_index_parse_fasta(const char *const *paths, uint64_t n_paths, kslam_index **out) {     // createIndexFromFASTA, :224-260
  if (!paths || !out) return KSLAM_ERR_ARG;
  kslam_index *index = nullptr;
  try {
    index = new kslam_index();
    for (uint64_t f = 0; f < n_paths; f++) {
      std::string data;
      if (!read_file(paths[f], data)) throw std::runtime_error(std::string("unable to open FASTA file ") + paths[f]);
      Entry cur;                                           // one object per file: bases before the first header form an entry too
      auto close = [&]() {
        if (cur.bases.size()) {
          for (char &ch : cur.bases) ch = (char)toupper((unsigned char)ch);   // inPlaceConvertToUpperCase
          index->add(cur);
        }
      };
      for (size_t pos = 0; pos < data.size();) {           // safeGetline (sequenceTools.h:45-73): "\n", "\r\n" and lone "\r" end a line
        size_t e = pos;
        while (e < data.size() && data[e] != '\n' && data[e] != '\r') e++;
        const char *line = data.data() + pos;
        const size_t len = e - pos;
        pos = e + ((e + 1 < data.size() && data[e] == '\r' && data[e + 1] == '\n') ? 2 : 1);
        if (len == 0) continue;
        if (line[0] == '>') {
          close();
          cur.bases.clear(); cur.locusTag.clear();
          const void *sp = memchr(line, ' ', len);
          if (sp && sp != line) cur.locusTag.assign(line + 1, (const char *)sp - line - 1);
        } else cur.bases.append(line, len);
      }
      close();
    }
    *out = index;
    return KSLAM_OK;
  } catch (const std::exception &e) { g_error = e.what(); delete index; return KSLAM_ERR_ARG; }
}

int kslam_index_read(const char *path, kslam_index **out) {           // getIndexFromBoostSerial, :336-344
  if (!path || !out) return KSLAM_ERR_ARG;
  kslam_index *index = nullptr;
  uint64_t src_size = 0, src_mtime = 0;
  const bool use_cache = !getenv("KSLAM_NO_INDEX_CACHE") && source_key(path, &src_size, &src_mtime);
  const std::string cache_path = std::string(path) + ".kslam";
  if (use_cache && (index = cache_load(cache_path, src_size, src_mtime)) != nullptr) { *out = index; return KSLAM_OK; }
  try {
    std::string data;
    if (!read_file(path, data)) throw std::runtime_error("unable to open index file");
    if (data.compare(0, sizeof kArchiveHeader - 1, kArchiveHeader) != 0) throw std::runtime_error("not a Boost text archive (header missing)");
    ArchiveReader r{data, sizeof kArchiveHeader - 1};
    if (r.number() < 4) throw std::runtime_error("archive library version < 4 is not supported");
    enum { INDEX = 1, ENTRIES = 2, ENTRY = 4, GENES = 8, GENE = 16, CDS = 32 };
    index = new kslam_index();
    index->bases.reserve(data.size());                     // the archive is almost all bases: no regrowth while appending
    r.class_info(INDEX);
    const uint64_t n_entries = r.vector(ENTRIES);
    for (uint64_t e = 0; e < n_entries; e++) {
      r.class_info(ENTRY);
      Entry en;
      const std::string_view bases = r.view();
      en.taxonomyID = (uint32_t)r.number(); en.genbankID = (uint32_t)r.number();
      en.isPlasmid = r.number() != 0; en.is16S = r.number() != 0;
      r.string(en.locusTag);
      const uint64_t n_genes = r.vector(GENES);
      for (uint64_t g = 0; g < n_genes; g++) {
        r.class_info(GENE);
        Gene ge;
        r.string(ge.geneName); r.string(ge.locusTag); r.string(ge.proteinID); r.string(ge.product); r.string(ge.referenceSequence);
        ge.geneID = (uint32_t)r.number();
        r.class_info(CDS);
        ge.start = (uint32_t)r.number(); ge.stop = (uint32_t)r.number(); ge.complement = r.number() != 0;
        en.genes.push_back(std::move(ge));
      }
      index->add(en, bases);
    }
    if (use_cache) cache_store(index, cache_path, src_size, src_mtime);
    *out = index;
    return KSLAM_OK;
  } catch (const std::exception &e) { g_error = e.what(); delete index; return KSLAM_ERR_ARG; }
}

int kslam_index_write(const kslam_index *ix, const char *path) {      // writeIndexToBoostSerial, :201-205
  if (!ix || !path) return KSLAM_ERR_ARG;
  FILE *f = fopen(path, "wb");
  if (!f) { g_error = "unable to open the database file for writing"; return KSLAM_ERR_ARG; }
  unsigned seen = 0;
  enum { INDEX = 1, ENTRIES = 2, ENTRY = 4, GENES = 8, GENE = 16, CDS = 32 };
  auto info = [&](unsigned bit) { if (!(seen & bit)) { seen |= bit; fputs(" 0 0", f); } };
  auto str = [&](const char *p, uint64_t n) { fprintf(f, " %llu ", (unsigned long long)n); fwrite(p, 1, n, f); };
  fprintf(f, "%s 17", kArchiveHeader);
  info(INDEX); info(ENTRIES);
  fprintf(f, " %llu 0", (unsigned long long)ix->n());
  for (uint64_t e = 0; e < ix->n(); e++) {
    info(ENTRY);
    str(ix->bases_data() + ix->offs[e], ix->offs[e + 1] - ix->offs[e]);
    fprintf(f, " %u %u %u %u", ix->taxonomy_ids[e], ix->genbank_ids[e], (unsigned)ix->is_plasmid[e], (unsigned)ix->is_16s[e]);
    str(ix->locus.data() + ix->locus_offs[e], ix->locus_offs[e + 1] - ix->locus_offs[e]);
    info(GENES);
    fprintf(f, " %llu 0", (unsigned long long)(ix->gene_offs[e + 1] - ix->gene_offs[e]));
    for (uint64_t g = ix->gene_offs[e]; g < ix->gene_offs[e + 1]; g++) {
      const kslam_gene &k = ix->genes[g];
      info(GENE);
      for (int i = 0; i < 5; i++) str(ix->gene_strings.data() + k.str_offs[i], k.str_offs[i + 1] - k.str_offs[i]);
      fprintf(f, " %u", k.gene_id);
      info(CDS);
      fprintf(f, " %u %u %u", k.cds_start, k.cds_stop, k.complement ? 1u : 0u);
    }
  }
  fputc('\n', f);
  const bool ok = !ferror(f);
  return (fclose(f) == 0 && ok) ? KSLAM_OK : KSLAM_ERR_STATE;
}

int kslam_gene_index_build(const kslam_sam_db *db, kslam_gene_index **out) {
  if (!db || !out || !db->genes || !db->gene_offs) return KSLAM_ERR_ARG;
  try {
    kslam_gene_index *gi = new kslam_gene_index();
    gi->sorted.assign(db->n_entries, 0);
    gi->max_stop.assign(db->gene_offs[db->n_entries], 0);
    for (uint64_t e = 0; e < db->n_entries; e++) {
      bool ok = true;
      int32_t run = INT32_MIN;
      for (uint64_t g = db->gene_offs[e]; g < db->gene_offs[e + 1]; g++) {
        const kslam_gene &k = db->genes[g];
        if (k.cds_start > (uint32_t)INT32_MAX || k.cds_stop > (uint32_t)INT32_MAX) ok = false;
        if (g > db->gene_offs[e] && k.cds_start < db->genes[g - 1].cds_start) ok = false;
        run = std::max(run, (int32_t)k.cds_stop);
        gi->max_stop[g] = run;
      }
      gi->sorted[e] = ok;
    }
    *out = gi;
    return KSLAM_OK;
  } catch (const std::exception &) { return KSLAM_ERR_NOMEM; }
}
void kslam_gene_index_free(kslam_gene_index *index) { delete index; }

int kslam_index_db(const kslam_index *ix, kslam_sam_db *out) {
  if (!ix || !out) return KSLAM_ERR_ARG;
  memset(out, 0, sizeof *out);
  out->n_entries = ix->n();
  out->bases = ix->bases_data(); out->offs = ix->offs.data();
  out->locus_tags = ix->locus.data(); out->locus_offs = ix->locus_offs.data();
  out->taxonomy_ids = ix->taxonomy_ids.data();
  if (!ix->genes.empty()) {
    out->genes = ix->genes.data(); out->gene_offs = ix->gene_offs.data(); out->gene_strings = ix->gene_strings.data();
    if (!ix->gene_index && kslam_gene_index_build(out, &ix->gene_index) != KSLAM_OK) ix->gene_index = nullptr;
    out->gene_index = ix->gene_index;
  }
  return KSLAM_OK;
}

}  // extern "C"
