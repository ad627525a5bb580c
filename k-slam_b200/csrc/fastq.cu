// fastq.cu — FASTQ ingest for the matching path (host code).
//
// Restates the reference's reader, /root/reference/src/FASTQsequence.h:129-165 (getSequencesFromFASTQFile,
// getPairedSequencesFromFASTQFiles :110-123) with its line reader sequenceTools.h:45-73 (safeGetline) and the read-id
// rule FASTQsequence.h:61-71, but chunk-parallel instead of one character at a time:
//   * The reference never validates a record: it counts lines, four per record (id, bases, '+' line ignored, quality).
//     A line ends at '\n', at "\r\n" or at a lone '\r'; a last line without terminator still counts when it is not
//     empty. So record boundaries are a function of line numbers only: every thread counts the line ends of its byte
//     range (a '\n' directly after a '\r' is absorbed), a prefix sum gives each range its first line number, and a
//     second pass writes the line start offsets. No guessing at '@' characters, which may start a quality line.
//   * One batch = the next `max_reads` records of R1 followed by the next `max_reads` records of R2 in one array (R1
//     block then R2 block), and the call fails with the reference's "mismatch in R1 and R2 size" rule
//     ((n1 + n2) / n1 != 2, integer division, FASTQsequence.h:117-122).
//   * Read id: drop the first character, cut at the first space, then at the first '/' (ids of length <= 1 are empty).
// The batch comes back in the layout kslam_align_batch / kslam_align_pair_batch take (one byte array + n+1 offsets),
// in page-locked memory when a CUDA device is present so the H2D copy runs at full PCIe rate.
#include "common.cuh"
#include <chrono>
#include <emmintrin.h>
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

namespace {

struct MappedFile {
  const char *p = nullptr;
  size_t size = 0, cursor = 0;
  bool eof_line_done = false;   // the empty line safeGetline yields once at end of file has been handed out
  double bytes_per_line = 0;    // running estimate from the last batch (sizes the next scan window)
  int fd = -1;
  char *heap = nullptr;         // contents of an input that cannot be mapped (pipe, process substitution, /dev/stdin)
  bool open(const char *path, std::string &err) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) { err = std::string("cannot open ") + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { err = std::string("cannot stat ") + path; return false; }
    if (!S_ISREG(st.st_mode)) {
      // The reference reads through std::ifstream, which works on FIFOs (`SLAM ... <(zcat r1.fq.gz)`); a pipe has no size
      // and cannot be mapped, so it is drained into memory and indexed like a mapping.
      size_t cap = 1u << 24;
      heap = (char *)malloc(cap);
      if (!heap) { err = "out of memory reading " + std::string(path); return false; }
      for (;;) {
        if (size == cap) {
          char *bigger = (char *)realloc(heap, cap * 2);
          if (!bigger) { err = "out of memory reading " + std::string(path); return false; }
          heap = bigger; cap *= 2;
        }
        const ssize_t got = ::read(fd, heap + size, cap - size);
        if (got < 0) { if (errno == EINTR) continue; err = std::string("read error on ") + path; return false; }
        if (got == 0) break;
        size += (size_t)got;
      }
      p = heap;
      return true;
    }
    size = (size_t)st.st_size;
    if (size) {
      void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) { err = std::string("cannot mmap ") + path; return false; }
      p = (const char *)m;
      madvise(m, size, MADV_SEQUENTIAL);
    }
    return true;
  }
  void close() {
    if (heap) free(heap);
    else if (p) munmap((void *)p, size);
    if (fd >= 0) ::close(fd);
    p = nullptr; heap = nullptr; fd = -1;
  }
};

// growable host buffer: page-locked when CUDA is usable, plain memory otherwise (the CPU test tier)
struct IngestBuf {
  char *p = nullptr; size_t cap = 0; bool pinned = false;
  void reserve(size_t bytes, bool want_pinned) {
    if (bytes <= cap) return;
    release();
    size_t want = bytes + bytes / 8 + 256;
    if (want_pinned && cudaMallocHost((void **)&p, want) == cudaSuccess) pinned = true;
    else { cudaGetLastError(); p = (char *)malloc(want); pinned = false; if (!p) throw std::bad_alloc(); }
    cap = want;
  }
  void release() { if (p) { if (pinned) cudaFreeHost(p); else free(p); } p = nullptr; cap = 0; }
};

template <class F> void parallel_for(uint32_t threads, uint64_t n, F f) {
  if (threads <= 1 || n < 2) { f(0, 0, n); return; }
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < threads; t++) {
    uint64_t lo = n * t / threads, hi = n * (t + 1) / threads;
    th.emplace_back([=] { f(t, lo, hi); });
  }
  for (auto &x : th) x.join();
}

// Offsets just past every '\n' of p[lo, hi) appended to `out` (the starts of the lines that follow). 64 bytes per step:
// four SSE2 compares folded into one 64-bit mask, one entry per set bit (memchr per line cost 4 calls per record); the same
// pass notices a '\r' anywhere in the range, which sends the window to the general path.
bool scan_newlines(const char *p, size_t lo, size_t hi, std::vector<uint64_t> &out) {   // returns true when the range holds a '\r'
  size_t i = lo;
  const __m128i nl = _mm_set1_epi8('\n'), cr = _mm_set1_epi8('\r');
  __m128i any_cr = _mm_setzero_si128();
  for (; i + 64 <= hi; i += 64) {
    const __m128i a = _mm_loadu_si128((const __m128i *)(p + i)), b = _mm_loadu_si128((const __m128i *)(p + i + 16));
    const __m128i c = _mm_loadu_si128((const __m128i *)(p + i + 32)), d = _mm_loadu_si128((const __m128i *)(p + i + 48));
    any_cr = _mm_or_si128(any_cr, _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(a, cr), _mm_cmpeq_epi8(b, cr)), _mm_or_si128(_mm_cmpeq_epi8(c, cr), _mm_cmpeq_epi8(d, cr))));
    uint64_t m = (uint64_t)(uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(a, nl)) | ((uint64_t)(uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(b, nl)) << 16) |
                 ((uint64_t)(uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(c, nl)) << 32) | ((uint64_t)(uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(d, nl)) << 48);
    while (m) { out.push_back(i + (size_t)__builtin_ctzll(m) + 1); m &= m - 1; }
  }
  bool tail_cr = false;
  for (; i < hi; i++) { if (p[i] == '\n') out.push_back(i + 1); tail_cr |= p[i] == '\r'; }
  return tail_cr || _mm_movemask_epi8(any_cr) != 0;
}

// is byte i the end of a line? ('\r', or a '\n' that does not directly follow a '\r')
inline bool line_end_at(const char *p, size_t i) { return p[i] == '\r' || (p[i] == '\n' && (i == 0 || p[i - 1] != '\r')); }

}  // namespace

struct kslam_fastq {
  MappedFile f[2];
  bool paired = false;
  uint32_t threads = 1;
  bool pinned = false;
  std::string err;
  // ring of batch buffers: batch b lives in set b % ring, so a caller that pipelines (ingest | GPU | SAM) can keep the
  // last ring - 1 batches alive without copying them (kslam_fastq_set_ring; default 1 = valid until the next call)
  struct BufSet { IngestBuf bases, ids, quals; std::vector<uint64_t> offs, id_offs, q_offs; };
  std::vector<BufSet> sets = std::vector<BufSet>(1);
  uint64_t n_batches = 0;
  std::vector<uint64_t> line_start[2];
  uint64_t n_file[2] = {0, 0};
};

// Finds the starts of the next 4 * max_reads lines of `mf` from its cursor (fewer at end of file). starts[k] = offset of
// line k, starts[n_lines] = offset just past the last line taken. Lines are taken in whole records only, except that —
// like the reference — a truncated last record is consumed and dropped.
static uint64_t index_lines(MappedFile &mf, uint64_t max_reads, uint32_t threads, std::vector<uint64_t> &starts, uint64_t *consumed_to) {
  const char *p = mf.p;
  const size_t begin = mf.cursor, size = mf.size;
  const uint64_t want_lines = 4 * max_reads;
  starts.clear();
  if (max_reads == 0 || (begin >= size && mf.eof_line_done)) { *consumed_to = begin; return 0; }
  if (begin >= size) {            // only the end-of-file empty line is left (see below)
    mf.eof_line_done = true;
    starts.assign(2, size);
    *consumed_to = size;
    return 1;
  }
  // grow a window until it holds want_lines line ends or reaches the end of the file
  size_t win = (size_t)(max_reads < (1u << 20) ? max_reads : (1u << 20)) * 512 + 4096;
  if (mf.bytes_per_line > 0) {                             // later batches: the file's own average line length, + 2 %
    const double est = (double)want_lines * mf.bytes_per_line * 1.02 + 65536;
    win = est < (double)(size - begin) ? (size_t)est : size - begin;
  }
  if (win > size - begin) win = size - begin;
  // Unix files (no '\r' in the window — the normal case): ONE scan. Every thread lists the line starts of its byte range,
  // the window grows by scanning only what is new, and the lists are joined in order. Files with '\r' take the general
  // path below (count, prefix sum, second pass).
  {
    std::vector<std::vector<uint64_t>> lists;              // in file order
    size_t scanned = begin, end_f = begin;
    uint64_t total_f = 0;
    bool cr_seen = false;
    for (;;) {
      end_f = begin + win;
      const size_t span = end_f - scanned;
      const uint32_t nt = (threads > 1 && span >= (1u << 16)) ? threads : 1;
      std::vector<uint8_t> cr(nt, 0);
      const size_t first = lists.size();
      lists.resize(first + nt);
      parallel_for(nt, span, [&](uint32_t t, uint64_t lo, uint64_t hi) {
        lists[first + t].reserve((size_t)((hi - lo) / 64) + 16);
        cr[t] = scan_newlines(p, scanned + lo, scanned + hi, lists[first + t]);
      });
      for (uint8_t v : cr) cr_seen |= v != 0;
      if (cr_seen) break;
      for (size_t l = first; l < lists.size(); l++) total_f += lists[l].size();
      scanned = end_f;
      if (total_f >= want_lines || end_f >= size) break;
      const double per_line = (double)win / (double)(total_f ? total_f : 1);
      const size_t need = (size_t)((double)(want_lines - total_f) * per_line * 1.25) + 65536;
      win = win + need > size - begin ? size - begin : win + need;
    }
    if (!cr_seen) {
      const bool at_eof = end_f >= size;
      const bool tail_line = at_eof && size > begin && p[size - 1] != '\n';   // a last line without terminator counts (it is not empty)
      uint64_t n_lines = total_f + (tail_line ? 1 : 0);
      bool eof_line = false;                               // the one EMPTY line safeGetline yields at end of file (see below)
      if (at_eof && n_lines < want_lines) { eof_line = true; n_lines++; mf.eof_line_done = true; }
      if (n_lines > want_lines) n_lines = want_lines;
      starts.assign(n_lines + 1, 0);
      starts[0] = begin;
      uint64_t k = 0;
      for (auto &l : lists) {
        if (k >= n_lines) break;
        const uint64_t take = std::min<uint64_t>(l.size(), n_lines - k);
        if (take) memcpy(&starts[k + 1], l.data(), take * sizeof(uint64_t));
        k += take;
      }
      if (tail_line && n_lines >= total_f + 1) starts[total_f + 1] = size;
      if (eof_line) starts[n_lines] = size;
      *consumed_to = starts[n_lines];
      if (n_lines > 64) mf.bytes_per_line = (double)(starts[n_lines] - begin) / (double)n_lines;
      return n_lines;
    }
    if (win > size - begin) win = size - begin;
  }
  std::vector<uint64_t> cnt(threads + 1);
  size_t end = begin;
  uint64_t total = 0;
  bool has_cr = false;      // Unix files (no '\r') take the vectorised path: count / find '\n' only
  for (;;) {
    end = begin + win;
    std::vector<uint8_t> cr(threads + 1, 0);
    parallel_for(threads, win, [&](uint32_t t, uint64_t lo, uint64_t hi) {
      cr[t] = memchr(p + begin + lo, '\r', hi - lo) != nullptr;
    });
    has_cr = false;
    for (uint8_t v : cr) has_cr |= v != 0;
    parallel_for(threads, win, [&](uint32_t t, uint64_t lo, uint64_t hi) {
      uint64_t c = 0;
      if (has_cr) for (size_t i = begin + lo; i < begin + hi; i++) c += line_end_at(p, i);
      else { const char *q = p + begin; for (size_t i = lo; i < hi; i++) c += q[i] == '\n'; }
      cnt[t + 1] = c;
    });
    cnt[0] = 0; total = 0;
    for (uint32_t t = 0; t < (threads > 1 && win >= 2 ? threads : 1); t++) { total += cnt[t + 1]; cnt[t + 1] = total; }
    if (total >= want_lines || end >= size) break;
    const double per_line = (double)win / (double)(total ? total : 1);
    size_t need = (size_t)((double)(want_lines - total) * per_line * 1.25) + 65536;
    win = win + need > size - begin ? size - begin : win + need;
  }
  const bool at_eof = end >= size;
  // a last line without terminator counts when it is not empty (sequenceTools.h:64-67)
  const bool tail_line = at_eof && size > begin && !line_end_at(p, size - 1) && !(p[size - 1] == '\n');
  uint64_t n_lines = total + (tail_line ? 1 : 0);
  // At end of file safeGetline returns one more, EMPTY line before the stream goes bad (it sets eofbit, which does not
  // make the stream false; sequenceTools.h:64-67): a file whose last record lacks its quality line still yields that
  // record, with an empty quality. Handed out once per file.
  bool eof_line = false;
  if (at_eof && n_lines < want_lines) { eof_line = true; n_lines++; mf.eof_line_done = true; }
  if (n_lines > want_lines) n_lines = want_lines;
  starts.assign(n_lines + 1, 0);
  // second pass: line k+1 starts right after the k-th line end (and after the '\n' a '\r' absorbs)
  const uint32_t nt = (threads > 1 && win >= 2) ? threads : 1;
  parallel_for(nt, win, [&](uint32_t t, uint64_t lo, uint64_t hi) {
    uint64_t k = cnt[t];
    if (!has_cr) {
      const char *q = p + begin + lo, *e = p + begin + hi;
      while (k < n_lines && q < e && (q = (const char *)memchr(q, '\n', (size_t)(e - q))) != nullptr) { q++; k++; starts[k] = (uint64_t)(q - p); }
      return;
    }
    for (size_t i = begin + lo; i < begin + hi && k < n_lines; i++)   // k counts real line ends; virtual lines are patched below
      if (line_end_at(p, i)) {
        size_t nxt = i + 1;
        if (p[i] == '\r' && nxt < size && p[nxt] == '\n') nxt++;
        k++;
        if (k <= n_lines) starts[k] = nxt;
      }
  });
  starts[0] = begin;
  if (tail_line && n_lines >= total + 1) starts[total + 1] = size;
  if (eof_line) starts[n_lines] = size;
  *consumed_to = starts[n_lines];
  return n_lines;
}

static inline size_t line_len(const char *p, uint64_t a, uint64_t b) {   // [a, b) minus its terminator
  size_t e = b;
  if (e > a && p[e - 1] == '\n') { e--; if (e > a && p[e - 1] == '\r') e--; }
  else if (e > a && p[e - 1] == '\r') e--;
  return e - a;
}

extern "C" {

int kslam_fastq_open(const char *r1_path, const char *r2_path, uint32_t threads, kslam_fastq **out) {
  if (!r1_path || !out) return KSLAM_ERR_ARG;
  *out = nullptr;
  kslam_fastq *rd = new kslam_fastq();
  rd->threads = threads ? threads : std::max(1u, std::thread::hardware_concurrency());
  if (rd->threads > 64) rd->threads = 64;
  rd->paired = r2_path != nullptr;
  int ndev = 0;
  rd->pinned = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
  cudaGetLastError();
  if (!rd->f[0].open(r1_path, rd->err) || (rd->paired && !rd->f[1].open(r2_path, rd->err))) {
    api_fail(nullptr, KSLAM_ERR_ARG, rd->err);
    rd->f[0].close(); rd->f[1].close();
    delete rd;
    return KSLAM_ERR_ARG;
  }
  *out = rd;
  return KSLAM_OK;
}

const char *kslam_fastq_error(const kslam_fastq *rd) { return rd ? rd->err.c_str() : ""; }

int kslam_fastq_set_ring(kslam_fastq *rd, uint32_t n_buffer_sets) {
  if (!rd || n_buffer_sets < 1 || n_buffer_sets > 16 || rd->n_batches) return KSLAM_ERR_ARG;
  rd->sets.resize(n_buffer_sets);
  return KSLAM_OK;
}

int kslam_fastq_next(kslam_fastq *rd, uint64_t max_reads, kslam_read_batch *out) {
  if (!rd || !out) return KSLAM_ERR_ARG;
  try {
    const bool trace = getenv("KSLAM_FASTQ_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    memset(out, 0, sizeof *out);
    const int nf = rd->paired ? 2 : 1;
    uint64_t n_rec[2] = {0, 0};
    for (int k = 0; k < nf; k++) {
      if (k == 1 && n_rec[0] == 0) break;                      // FASTQsequence.h:116: R2 is not touched when R1 is empty
      uint64_t to = rd->f[k].cursor;
      const uint64_t n_lines = index_lines(rd->f[k], max_reads, rd->threads, rd->line_start[k], &to);
      rd->f[k].cursor = to;
      n_rec[k] = n_lines / 4;                                  // a truncated last record is consumed and dropped
    }
    const uint64_t n = n_rec[0] + n_rec[1];
    if (rd->paired && n_rec[0] && n / n_rec[0] != 2) {         // FASTQsequence.h:117-122 (integer division, as written)
      rd->err = "mismatch in R1 and R2 size";
      return KSLAM_ERR_STATE;
    }
    const double t1 = now();
    kslam_fastq::BufSet &bs = rd->sets[rd->n_batches % rd->sets.size()];
    rd->n_batches++;
    bs.offs.assign(n + 1, 0); bs.id_offs.assign(n + 1, 0); bs.q_offs.assign(n + 1, 0);
    std::vector<uint64_t> &offs = bs.offs, &ioffs = bs.id_offs, &qoffs = bs.q_offs;
    // lengths, then prefix sums (ids are parsed twice: once for the length, once for the copy)
    auto rec = [&](uint64_t i, int *file, uint64_t *r) { if (i < n_rec[0]) { *file = 0; *r = i; } else { *file = 1; *r = i - n_rec[0]; } };
    auto id_span = [&](const char *p, uint64_t a, size_t len, size_t *from, size_t *cnt) {
      *from = 0; *cnt = 0;
      if (len <= 1) return;                                    // FASTQsequence.h:64
      const char *sp = (const char *)memchr(p + a, ' ', len);
      size_t c = sp ? (size_t)(sp - (p + a)) : len;            // substr(1, spacePos - 1) / npos - 1
      if (c == 0) return;                                      // space at position 0: substr(1, 0)
      c -= 1;
      const char *sl = (const char *)memchr(p + a + 1, '/', c);
      if (sl) c = (size_t)(sl - (p + a + 1));
      *from = 1; *cnt = c;
    };
    parallel_for(rd->threads, n, [&](uint32_t, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; i++) {
        int fl; uint64_t r; rec(i, &fl, &r);
        const char *p = rd->f[fl].p; const std::vector<uint64_t> &ls = rd->line_start[fl];
        offs[i + 1] = line_len(p, ls[4 * r + 1], ls[4 * r + 2]);
        qoffs[i + 1] = line_len(p, ls[4 * r + 3], ls[4 * r + 4]);
        size_t from, cnt; id_span(p, ls[4 * r], line_len(p, ls[4 * r], ls[4 * r + 1]), &from, &cnt);
        ioffs[i + 1] = cnt;
      }
    });
    const double t2 = now();
    for (uint64_t i = 0; i < n; i++) { offs[i + 1] += offs[i]; ioffs[i + 1] += ioffs[i]; qoffs[i + 1] += qoffs[i]; }
    const double t3 = now();
    bs.bases.reserve(offs[n] + 16, rd->pinned); bs.quals.reserve(qoffs[n] + 16, false); bs.ids.reserve(ioffs[n] + 16, false);
    parallel_for(rd->threads, n, [&](uint32_t, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; i++) {
        int fl; uint64_t r; rec(i, &fl, &r);
        const char *p = rd->f[fl].p; const std::vector<uint64_t> &ls = rd->line_start[fl];
        const size_t bl = offs[i + 1] - offs[i];
        memcpy(bs.bases.p + offs[i], p + ls[4 * r + 1], bl);
        memcpy(bs.quals.p + qoffs[i], p + ls[4 * r + 3], qoffs[i + 1] - qoffs[i]);     // stored as is, whatever its length
        size_t from, cnt; id_span(p, ls[4 * r], line_len(p, ls[4 * r], ls[4 * r + 1]), &from, &cnt);
        memcpy(bs.ids.p + ioffs[i], p + ls[4 * r] + from, cnt);
      }
    });
    if (trace) fprintf(stderr, "[kslam_fastq] %llu reads: line index %.1f ms, lengths %.1f ms, prefix sums %.1f ms, copy %.1f ms\n",
                       (unsigned long long)n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (now() - t3) * 1e3);
    out->n_reads = n; out->n_r1 = n_rec[0];
    out->bases = bs.bases.p; out->offs = bs.offs.data();
    out->quals = bs.quals.p; out->qual_offs = bs.q_offs.data();
    out->ids = bs.ids.p; out->id_offs = bs.id_offs.data();
    rd->n_file[0] += n_rec[0]; rd->n_file[1] += n_rec[1];
    return KSLAM_OK;
  } catch (const std::exception &e) { rd->err = e.what(); return KSLAM_ERR_NOMEM; }
}

void kslam_fastq_close(kslam_fastq *rd) {
  if (!rd) return;
  rd->f[0].close(); rd->f[1].close();
  for (auto &b : rd->sets) { b.bases.release(); b.ids.release(); b.quals.release(); }
  delete rd;
}

}  // extern "C"
