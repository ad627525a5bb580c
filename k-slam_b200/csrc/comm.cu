// comm.cu — the k-mer-range partitioned matching path driven from the C++ host over NCCL (BASELINE config 4).
//
// dist.cu holds the per-rank stages (kslam_part_route_kmers / _join / _finish); this file is the protocol around them, one
// exchange each way, issued as ncclSend / ncclRecv groups on the ctx's own stream over NVLink / NVSwitch:
//
//   all ranks     ncclAllGather      read counts -> first job-global read id of every rank (KMer.h:65-66: 30 bits)
//   read owner    route              extract + prefilter read k-mers, bucket by key owner
//   all ranks     ncclAllGather      bucket sizes (n_ranks^2 counts)
//   all ranks     ncclSend/Recv      16-byte k-mer records to the owner of their key range
//   key owner     join               radix sort, merge-join against the local slice, bucket raw matches by read owner
//   all ranks     ncclAllGather + ncclSend/Recv   16-byte match records back to the read owner
//   read owner    finish             match -> seed, seed sort, fuzzy unique, Smith-Waterman: the single-GPU path unchanged
//
// Equal k-mers share an owner, so no pile (Overlap.h:153-199) is split — the invariant of the reference's own chunking
// (Overlap.h:285-287) — and every rank's result is bit-identical to kslam_align_batch on its reads.
//
// NCCL is opened at run time (dlopen "libnccl.so.2"): a single-GPU user of libkslam.so never needs it, and inside a process
// that already carries a NCCL (torch) the same library is reused. Two ways to form the communicator:
//   kslam_comm_init_all   one process, one ctx per device (what `SLAM --devices a,b,..` uses; ncclCommInitAll);
//   kslam_comm_unique_id + kslam_comm_init_rank   one process per GPU (torchrun / MPI launchers distribute the id).
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi *nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) { api.error = "NCCL not found (dlopen libnccl.so.2)"; return; }
#define KS_SYM(field, sym) \
    *(void **)(&api.field) = dlsym(api.lib, sym); \
    if (!api.field) { api.error = std::string("NCCL symbol missing: ") + sym; return; }
    KS_SYM(GetUniqueId, "ncclGetUniqueId") KS_SYM(CommInitRank, "ncclCommInitRank") KS_SYM(CommInitAll, "ncclCommInitAll")
    KS_SYM(CommDestroy, "ncclCommDestroy") KS_SYM(AllGather, "ncclAllGather") KS_SYM(Send, "ncclSend") KS_SYM(Recv, "ncclRecv")
    KS_SYM(GroupStart, "ncclGroupStart") KS_SYM(GroupEnd, "ncclGroupEnd") KS_SYM(GetErrorString, "ncclGetErrorString")
#undef KS_SYM
  });
  return &api;
}

struct NcclError { ncclResult_t r; const char *what; };
#define NCCL_TRY(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) throw NcclError{r_, #x}; } while (0)

}  // namespace

struct kslam_comm {
  kslam_ctx *ctx = nullptr;
  ncclComm_t comm = nullptr;
  uint32_t rank = 0, n_ranks = 1;
  DevBuf d_small;              // u64[(n_ranks + 1) * n_ranks]: a row to send + the gathered matrix
  HostBuf h_small;
  kslam_comm_stats st{};
};

namespace {

int comm_fail(kslam_comm *m, int code, const std::string &msg) { return api_fail(m ? m->ctx : nullptr, code, msg); }

// every rank contributes `n` u64 values; all[r * n + i] = value i of rank r (host, after the call)
void allgather_u64(kslam_comm *m, const uint64_t *mine, uint32_t n, uint64_t *all) {
  kslam_ctx *c = m->ctx;
  const size_t row = (size_t)n * 8;
  m->d_small.reserve(row * (m->n_ranks + 1) + 64); m->h_small.reserve(row * (m->n_ranks + 1) + 64);
  uint64_t *h = m->h_small.as<uint64_t>();
  for (uint32_t i = 0; i < n; i++) h[i] = mine[i];
  char *d = m->d_small.as<char>();
  CUDA_TRY(cudaMemcpyAsync(d, h, row, cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(nccl()->AllGather(d, d + row, n, ncclUint64, m->comm, c->stream));
  CUDA_TRY(cudaMemcpyAsync(h + n, d + row, row * m->n_ranks, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < (size_t)n * m->n_ranks; i++) all[i] = h[n + i];
}

// A collective must be entered by every rank or the others wait for ever, so a rank whose local stage failed still takes
// part in the size exchange that follows it, with a flag next to its (zero) counts; every rank then sees the flag and
// leaves the batch before the data exchange. Thrown by the ranks that did NOT fail themselves.
struct PeerFailed { uint32_t rank; };

// counts[p] records of `send` (grouped by destination, in rank order) go to rank p; recv is filled in source-rank order.
// Returns the number of records received; *ms = device time of the exchange. failed: this rank has nothing to send
// because its stage failed — returns ~0 on it, throws PeerFailed on the others, and nobody exchanges records.
uint64_t all_to_all_records(kslam_comm *m, const Rec16 *send, const uint64_t *counts, bool failed, bool matches, float *ms, uint64_t *bytes_out) {
  kslam_ctx *c = m->ctx;
  const uint32_t P = m->n_ranks;
  std::vector<uint64_t> row(P + 1, 0), gathered((size_t)(P + 1) * P), all((size_t)P * P);
  if (!failed) for (uint32_t p = 0; p < P; p++) row[p] = counts[p];
  row[P] = failed ? 1 : 0;
  allgather_u64(m, row.data(), P + 1, gathered.data());
  for (uint32_t src = 0; src < P; src++) {
    if (gathered[(size_t)src * (P + 1) + P]) { if (failed) return ~0ull; throw PeerFailed{src}; }
    for (uint32_t p = 0; p < P; p++) all[(size_t)src * P + p] = gathered[(size_t)src * (P + 1) + p];
  }
  uint64_t n_recv = 0;
  for (uint32_t src = 0; src < P; src++) n_recv += all[(size_t)src * P + m->rank];
  void *recv_ptr = nullptr;
  if ((matches ? kslam_part_match_buffer(c, n_recv, &recv_ptr) : kslam_part_recv_buffer(c, n_recv, &recv_ptr)) != KSLAM_OK)
    throw ArgError{"could not reserve the receive buffer of the exchange"};
  Rec16 *recv = (Rec16 *)recv_ptr;
  cudaEvent_t e0 = tm_mark(c);
  NCCL_TRY(nccl()->GroupStart());
  uint64_t soff = 0, roff = 0, sent = 0;
  for (uint32_t p = 0; p < P; p++) {
    const uint64_t ns = counts[p], nr = all[(size_t)p * P + m->rank];
    if (ns) NCCL_TRY(nccl()->Send(send + soff, ns * sizeof(Rec16), ncclUint8, (int)p, m->comm, c->stream));
    if (nr) NCCL_TRY(nccl()->Recv(recv + roff, nr * sizeof(Rec16), ncclUint8, (int)p, m->comm, c->stream));
    if (p != m->rank) sent += ns;
    soff += ns; roff += nr;
  }
  NCCL_TRY(nccl()->GroupEnd());
  cudaEvent_t e1 = tm_mark(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  *ms = tm_ms(e0, e1);
  *bytes_out = sent * sizeof(Rec16);
  return n_recv;
}

}  // namespace

#define COMM_BEGIN(m)                                                                          \
  if (!(m) || !(m)->ctx) return KSLAM_ERR_ARG;                                                 \
  kslam_ctx *c = (m)->ctx;                                                                     \
  API_BEGIN(c)
#define COMM_END(m)                                                                            \
  } catch (const PeerFailed &e) {                                                              \
    return api_fail(c, KSLAM_ERR_STATE, "rank " + std::to_string(e.rank) + " of the communicator failed in this batch (its own error says why)"); \
  } catch (const NcclError &e) {                                                               \
    return api_fail(c, KSLAM_ERR_CUDA, std::string(e.what) + ": " + nccl()->GetErrorString(e.r)); \
  API_END(c)

extern "C" {

int kslam_comm_unique_id(void *id128) {
  if (!id128) return KSLAM_ERR_ARG;
  if (!nccl()->error.empty()) return api_fail(nullptr, KSLAM_ERR_STATE, nccl()->error);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (nccl()->GetUniqueId(&id) != ncclSuccess) return api_fail(nullptr, KSLAM_ERR_CUDA, "ncclGetUniqueId failed");
  memcpy(id128, &id, sizeof id);
  return KSLAM_OK;
}

int kslam_comm_init_rank(kslam_ctx *ctx, uint32_t rank, uint32_t n_ranks, const void *id128, kslam_comm **out) {
  if (!ctx || !out || !id128 || rank >= n_ranks || n_ranks > 64) return KSLAM_ERR_ARG;
  *out = nullptr;
  if (!nccl()->error.empty()) return api_fail(ctx, KSLAM_ERR_STATE, nccl()->error);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return api_fail(ctx, KSLAM_ERR_CUDA, "cudaSetDevice failed");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t comm = nullptr;
  const ncclResult_t r = nccl()->CommInitRank(&comm, (int)n_ranks, id, (int)rank);
  if (r != ncclSuccess) return api_fail(ctx, KSLAM_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl()->GetErrorString(r));
  kslam_comm *m = new kslam_comm();
  m->ctx = ctx; m->comm = comm; m->rank = rank; m->n_ranks = n_ranks;
  *out = m;
  return KSLAM_OK;
}

int kslam_comm_init_all(uint32_t n, kslam_ctx *const *ctxs, kslam_comm **out) {
  if (!n || n > 64 || !ctxs || !out) return KSLAM_ERR_ARG;
  for (uint32_t i = 0; i < n; i++) { out[i] = nullptr; if (!ctxs[i]) return KSLAM_ERR_ARG; }
  if (!nccl()->error.empty()) return api_fail(ctxs[0], KSLAM_ERR_STATE, nccl()->error);
  std::vector<int> devs(n);
  for (uint32_t i = 0; i < n; i++) devs[i] = ctxs[i]->device;
  std::vector<ncclComm_t> comms(n, nullptr);
  const ncclResult_t r = nccl()->CommInitAll(comms.data(), (int)n, devs.data());
  if (r != ncclSuccess) return api_fail(ctxs[0], KSLAM_ERR_CUDA, std::string("ncclCommInitAll (one ctx per DISTINCT device): ") + nccl()->GetErrorString(r));
  for (uint32_t i = 0; i < n; i++) {
    kslam_comm *m = new kslam_comm();
    m->ctx = ctxs[i]; m->comm = comms[i]; m->rank = i; m->n_ranks = n;
    out[i] = m;
  }
  return KSLAM_OK;
}

void kslam_comm_destroy(kslam_comm *m) {
  if (!m) return;
  if (m->ctx) cudaSetDevice(m->ctx->device);
  if (m->comm && nccl()->CommDestroy) nccl()->CommDestroy(m->comm);
  m->d_small.release(); m->h_small.release();
  delete m;
}

int kslam_comm_rank(const kslam_comm *m, uint32_t *rank, uint32_t *n_ranks) {
  if (!m) return KSLAM_ERR_ARG;
  if (rank) *rank = m->rank;
  if (n_ranks) *n_ranks = m->n_ranks;
  return KSLAM_OK;
}

// alignToDatabase over the partitioned index for the reads this rank has uploaded (kslam_upload_reads). Collective: every
// rank of the communicator calls it once per batch (a rank without reads uploads an empty batch).
// Every rank of the communicator enters the batch (first_gather) whatever state it is in: a rank that cannot take part
// says so there, and all ranks return an error instead of waiting for it.
static void first_gather(kslam_comm *m, uint64_t n_reads_mine, bool failed, std::vector<uint64_t> &n_reads) {
  const uint32_t P = m->n_ranks;
  const uint64_t mine[2] = {failed ? 0 : n_reads_mine, failed ? 1ull : 0ull};
  std::vector<uint64_t> g((size_t)2 * P);
  allgather_u64(m, mine, 2, g.data());
  n_reads.resize(P);
  for (uint32_t p = 0; p < P; p++) n_reads[p] = g[(size_t)2 * p];
  if (failed) return;
  for (uint32_t p = 0; p < P; p++) if (g[(size_t)2 * p + 1]) throw PeerFailed{p};
}

int kslam_comm_abort_batch(kslam_comm *m) {
  COMM_BEGIN(m)
  std::vector<uint64_t> n_reads;
  first_gather(m, 0, true, n_reads);
  return KSLAM_OK;
  COMM_END(m)
}

int kslam_comm_align_resident(kslam_comm *m, int fetch, kslam_alignments *out) {
  COMM_BEGIN(m)
  const uint32_t P = m->n_ranks;
  kslam_comm_stats &st = m->st;
  memset(&st, 0, sizeof st);
  const char *not_ready = !c->reads_loaded ? "kslam_upload_reads first"
                          : (c->n_parts != m->n_ranks || c->part != m->rank) ? "kslam_load_genomes_part(part = rank, n_parts = ranks of the communicator) first" : nullptr;
  // job-global read ids: rank r's reads are numbered from the sum of the counts before it
  std::vector<uint64_t> n_reads;
  first_gather(m, c->reads.n, not_ready != nullptr, n_reads);
  if (not_ready) return comm_fail(m, KSLAM_ERR_STATE, not_ready);
  std::vector<uint32_t> id_bases(P + 1, 0);
  uint64_t run = 0;
  for (uint32_t p = 0; p < P; p++) { id_bases[p] = (uint32_t)run; run += n_reads[p]; }
  if (run > (1ull << 30)) return comm_fail(m, KSLAM_ERR_ARG, "more than 2^30 reads in one job-wide batch (KMer.h:65-66)");   // (the same sum on every rank)
  id_bases[P] = (uint32_t)run;
  // read owner: extract, prefilter, bucket by key owner
  const void *send = nullptr;
  std::vector<uint64_t> counts(P);
  int rc = kslam_part_route_kmers(c, id_bases[m->rank], &send, counts.data());
  st.ms_route = c->tm.ms_extract; st.ms_bucket_kmers = c->part_ms_bucket;
  if (rc == KSLAM_OK) for (uint32_t p = 0; p < P; p++) st.kmers_sent += counts[p];
  const uint64_t n_recv = all_to_all_records(m, (const Rec16 *)send, counts.data(), rc != KSLAM_OK, false, &st.ms_exchange_kmers, &st.bytes_sent_kmers);
  if (rc != KSLAM_OK) return rc;                            // (the ctx's error text is the stage's)
  st.kmers_received = n_recv;
  // key owner: sort, merge-join, bucket raw matches by read owner
  const void *msend = nullptr;
  std::vector<uint64_t> mcounts(P);
  rc = kslam_part_join(c, n_recv, id_bases.data(), &msend, mcounts.data());
  st.ms_sort = c->tm.ms_sort; st.ms_join = c->tm.ms_join; st.ms_bucket_matches = c->part_ms_bucket_matches;
  if (rc == KSLAM_OK) for (uint32_t p = 0; p < P; p++) st.matches_sent += mcounts[p];
  const uint64_t n_m = all_to_all_records(m, (const Rec16 *)msend, mcounts.data(), rc != KSLAM_OK, true, &st.ms_exchange_matches, &st.bytes_sent_matches);
  if (rc != KSLAM_OK) return rc;
  st.matches_received = n_m;
  // read owner: the single-GPU path from the seeds on
  rc = kslam_part_finish(c, n_m, id_bases[m->rank], fetch, out);
  st.ms_finish = c->tm.ms_total;
  return rc;
  COMM_END(m)
}

int kslam_comm_get_stats(const kslam_comm *m, kslam_comm_stats *out) {
  if (!m || !out) return KSLAM_ERR_ARG;
  *out = m->st;
  return KSLAM_OK;
}

}  // extern "C"
