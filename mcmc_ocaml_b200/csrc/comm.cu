// comm.cu -- the multi-GPU side of the C ABI: one process per GPU, one NCCL communicator per context.
//
// What shards: chains and queries are independent (no data-path collective, chain ids are global);
// a kd-tree is built on one rank and replicated with ONE ncclBroadcast of its contiguous device blob, straight from
// the builder's blob into the receivers' blobs (no staging copy); per-rank statistics and per-cell evidence terms are
// all-gathered and combined in a fixed order, so every rank holds the same bits and the result does not depend on the
// number of ranks.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, a process that
// already carries an NCCL (torch) shares that copy, and without NCCL only these entry points fail (MG_EFAIL).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kdtree.cuh"

static_assert(sizeof(ncclUniqueId) == MG_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");

namespace mg {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  std::string why;
};

static NcclApi *nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    const char *names[] = {getenv("MCMC_GPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n || !*n) continue;
      a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.h) break;
    }
    if (!a.h) { a.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"); return a; }
#define MG_SYM(field, name)                                                    \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.h, name));              \
  if (!a.field) { a.why = std::string("libnccl lacks ") + name; a.h = nullptr; return a; }
    MG_SYM(GetUniqueId, "ncclGetUniqueId") MG_SYM(CommInitRank, "ncclCommInitRank") MG_SYM(CommDestroy, "ncclCommDestroy")
    MG_SYM(Broadcast, "ncclBroadcast") MG_SYM(AllGather, "ncclAllGather") MG_SYM(GetErrorString, "ncclGetErrorString")
    MG_SYM(GetVersion, "ncclGetVersion")
#undef MG_SYM
    return a;
  }();
  return &api;
}

}  // namespace mg

struct mg_comm {
  mg_ctx *ctx = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  double last_collective_ms = 0.0;
};

using namespace mg;

#define MG_NCCL(ctx, call)                                                                              \
  do {                                                                                                  \
    ncclResult_t r_ = (call);                                                                           \
    if (r_ != ncclSuccess)                                                                              \
      return mg::set_err((ctx), MG_EFAIL, "nccl: %s at %s:%d", nccl_api()->GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

extern "C" int mg_comm_get_unique_id(uint8_t id[MG_COMM_ID_BYTES]) {
  NcclApi *api = nccl_api();
  if (!api->h || !id) return MG_EFAIL;
  ncclUniqueId u;
  if (api->GetUniqueId(&u) != ncclSuccess) return MG_EFAIL;
  memcpy(id, &u, MG_COMM_ID_BYTES);
  return MG_OK;
}

extern "C" int mg_comm_create(mg_ctx *ctx, int32_t nranks, int32_t rank, const uint8_t id[MG_COMM_ID_BYTES], mg_comm **out) {
  if (!ctx) return MG_EINVAL;
  MG_REQUIRE(ctx, out && nranks >= 1 && rank >= 0 && rank < nranks, "comm_create: bad rank / size");
  *out = nullptr;
  mg_comm *c = new mg_comm;
  c->ctx = ctx; c->nranks = nranks; c->rank = rank;
  if (nranks > 1) {
    NcclApi *api = nccl_api();
    if (!api->h) { delete c; return set_err(ctx, MG_EFAIL, "nccl: %s", api->why.c_str()); }
    if (!id) { delete c; return set_err(ctx, MG_EINVAL, "comm_create: null unique id"); }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { delete c; return set_err(ctx, MG_ECUDA, "cuda: cannot select the device"); }
    ncclUniqueId u;
    memcpy(&u, id, MG_COMM_ID_BYTES);
    ncclResult_t r = api->CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) { delete c; return set_err(ctx, MG_EFAIL, "nccl: %s (ncclCommInitRank)", api->GetErrorString(r)); }
  }
  *out = c;
  return MG_OK;
}

extern "C" void mg_comm_destroy(mg_comm *c) {
  if (!c) return;
  if (c->comm) {
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    nccl_api()->CommDestroy(c->comm);
  }
  delete c;
}
extern "C" int32_t mg_comm_rank(const mg_comm *c) { return c ? c->rank : -1; }
extern "C" int32_t mg_comm_size(const mg_comm *c) { return c ? c->nranks : 0; }
extern "C" double mg_comm_last_collective_ms(const mg_comm *c) { return c ? c->last_collective_ms : 0.0; }
extern "C" int mg_comm_nccl_version(int32_t *version) {
  NcclApi *api = nccl_api();
  if (!api->h || !version) return MG_EFAIL;
  int v = 0;
  if (api->GetVersion(&v) != ncclSuccess) return MG_EFAIL;
  *version = v;
  return MG_OK;
}

namespace mg {

// device-to-device collectives on the context's stream, timed with events (the time a rank spends inside the call)
struct CollTimer {
  mg_comm *c; cudaEvent_t e0 = nullptr, e1 = nullptr;
  explicit CollTimer(mg_comm *c_) : c(c_) {
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, c->ctx->stream);
  }
  void stop() {
    cudaEventRecord(e1, c->ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) c->last_collective_ms = ms;
  }
  ~CollTimer() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

int comm_broadcast_dev(mg_comm *c, void *d_buf, size_t nbytes, int root) {
  if (c->nranks == 1 || nbytes == 0) return MG_OK;
  MG_NCCL(c->ctx, nccl_api()->Broadcast(d_buf, d_buf, nbytes, ncclUint8, root, c->comm, c->ctx->stream));
  c->ctx->launches++;
  return MG_OK;
}

int comm_allgather_dev(mg_comm *c, const void *d_send, void *d_recv, size_t nbytes_per_rank) {
  mg_ctx *ctx = c->ctx;
  if (c->nranks == 1) {
    if (d_send != d_recv) MG_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, nbytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
    return MG_OK;
  }
  MG_NCCL(ctx, nccl_api()->AllGather(d_send, d_recv, nbytes_per_rank, ncclUint8, c->comm, ctx->stream));
  ctx->launches++;
  return MG_OK;
}

}  // namespace mg

// All-gather of a small host record from every rank (per-rank counts, sums, partial statistics): recv holds
// nranks * nbytes in rank order on every rank.
extern "C" int mg_comm_allgather(mg_comm *c, const void *send, void *recv, int64_t nbytes) {
  if (!c) return MG_EINVAL;
  mg_ctx *ctx = c->ctx;
  MG_REQUIRE(ctx, send && recv && nbytes >= 0, "allgather: bad arguments");
  if (nbytes == 0) return MG_OK;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (c->nranks == 1) { memmove(recv, send, (size_t)nbytes); return MG_OK; }
  DevBuf<uint8_t> d_send, d_recv;
  MG_CUDA(ctx, d_send.alloc((size_t)nbytes, ctx->stream));
  MG_CUDA(ctx, d_recv.alloc((size_t)nbytes * c->nranks, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(d_send.get(), send, (size_t)nbytes, cudaMemcpyHostToDevice, ctx->stream));
  CollTimer tm(c);
  int rc = comm_allgather_dev(c, d_send.get(), d_recv.get(), (size_t)nbytes);
  if (rc) return rc;
  tm.stop();
  MG_CUDA(ctx, cudaMemcpyAsync(recv, d_recv.get(), (size_t)nbytes * c->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MG_OK;
}

extern "C" int mg_comm_barrier(mg_comm *c) {
  if (!c) return MG_EINVAL;
  uint8_t one = 1;
  std::vector<uint8_t> all((size_t)c->nranks);
  return mg_comm_allgather(c, &one, all.data(), 1);
}

// Kd-tree built on `root`, replicated on every rank: the header first (sizes), then ONE broadcast of the blob from the
// builder's blob straight into each receiver's freshly allocated blob.  On the root *out is `tree` itself.
extern "C" int mg_kdtree_broadcast(mg_comm *c, mg_kdtree *tree, int32_t root, mg_kdtree **out) {
  if (!c) return MG_EINVAL;
  mg_ctx *ctx = c->ctx;
  MG_REQUIRE(ctx, out && root >= 0 && root < c->nranks, "kdtree_broadcast: bad arguments");
  MG_REQUIRE(ctx, c->rank != root || tree != nullptr, "kdtree_broadcast: the root has no tree");
  MG_REQUIRE(ctx, c->rank != root || tree->ctx == ctx, "kdtree_broadcast: the tree belongs to another context");
  *out = nullptr;
  if (c->nranks == 1) { *out = tree; return MG_OK; }
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  int rc;
  DevBuf<uint8_t> d_hdr;
  constexpr size_t HB = 256;                     // the header occupies the blob's first 256-byte section
  static_assert(sizeof(KdHeader) <= HB, "header section");
  void *hdr_buf = nullptr;
  if (c->rank == root) hdr_buf = tree->d_blob;
  else { MG_CUDA(ctx, d_hdr.alloc(HB, s)); hdr_buf = d_hdr.get(); }
  CollTimer tm(c);
  if ((rc = comm_broadcast_dev(c, hdr_buf, HB, root))) return rc;
  if (c->rank == root) {
    if ((rc = comm_broadcast_dev(c, tree->d_blob, (size_t)tree->h.nbytes, root))) return rc;
    tm.stop();
    *out = tree;
    return MG_OK;
  }
  KdHeader h;
  MG_CUDA(ctx, cudaMemcpyAsync(&h, hdr_buf, sizeof h, cudaMemcpyDeviceToHost, s));
  MG_CUDA(ctx, cudaStreamSynchronize(s));
  // a bad header must not leave this rank out of the collective the others have entered: receive into a blob of the
  // announced size if that is at all possible, then judge
  const bool sane = h.magic == KD_MAGIC && h.nbytes >= (int64_t)HB && h.nbytes < (1LL << 40);
  if (!sane) return set_err(ctx, MG_EFAIL, "kdtree_broadcast: the root sent no kd-tree header");
  mg_kdtree *t = new mg_kdtree;
  t->ctx = ctx; t->h = h;
  cudaError_t e = cudaMallocAsync(&t->d_blob, (size_t)h.nbytes, s);
  if (e != cudaSuccess) { delete t; return set_err(ctx, MG_ENOMEM, "cuda: %s (kd-tree blob of %lld bytes)", cudaGetErrorString(e), (long long)h.nbytes); }
  rc = comm_broadcast_dev(c, t->d_blob, (size_t)h.nbytes, root);
  if (rc == MG_OK) { tm.stop(); rc = validate_blob_header(ctx, h); }
  if (rc) { cudaFreeAsync(t->d_blob, s); delete t; return rc; }
  *out = t;
  return MG_OK;
}

mg_ctx *mg_comm_ctx(const mg_comm *c) { return c ? c->ctx : nullptr; }

#include "reduce_sum.cuh"

// Evidence.evidence_harmonic_mean (evidence.ml:101-107) over samples SHARDED across ranks: every rank reduces the
// 1/L of its own shard (compensated), the (count, sum) pairs are all-gathered and folded in rank order.  Agrees with
// the single-GPU value to the rounding of the compensated sums (<= 2 ulp); identical on every rank.
extern "C" int mg_evidence_harmonic_mean_sharded(mg_comm *c, const double *d_ll_shard, int64_t n_shard, double *out) {
  if (!c) return MG_EINVAL;
  mg_ctx *ctx = c->ctx;
  MG_REQUIRE(ctx, out && n_shard >= 0 && (d_ll_shard || n_shard == 0), "evidence_harmonic_mean (sharded): bad arguments");
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  double mine[2] = {(double)n_shard, 0.0};
  if (n_shard > 0) {
    int rc = reduce_sum(ctx, n_shard, [d_ll_shard] __device__(int64_t i) { return 1.0 / exp(d_ll_shard[i]); }, &mine[1]);
    if (rc) return rc;
  }
  std::vector<double> all((size_t)2 * c->nranks);
  int rc = mg_comm_allgather(c, mine, all.data(), sizeof mine);
  if (rc) return rc;
  Comp linv; double n = 0.0;
  for (int r = 0; r < c->nranks; ++r) { n += all[2 * r]; linv.add(all[2 * r + 1]); }
  MG_REQUIRE(ctx, n >= 1.0, "evidence_harmonic_mean (sharded): no samples on any rank");
  *out = n / linv.value();
  return MG_OK;
}

// Mcmc.rjmcmc_array (mcmc.ml:121-139) with the ensemble's chains cut into contiguous ranges, one per rank: rank r runs
// the global chain ids cfg->chain_offset + [b_r, e_r) (the Philox stream is addressed by the global id, so the
// chains do not depend on the number of ranks), then rjmcmc_model_counts (mcmc.ml:141-149) and the accept / jump
// counters of all ranks are all-gathered and summed.  out_model / out_samples receive this rank's shard only
// ([n][C_r], [n][Dmax+2][C_r]); *shard_begin / *shard_count tell which chains those are.
extern "C" int mg_rjmcmc_array_sharded(mg_comm *c, const mg_rj_model *A, const mg_rj_model *B, const mg_rjmcmc_cfg *cfg,
                                       const double *a0, const double *b0, uint8_t *out_model, double *out_samples,
                                       int64_t out_counts[2], int64_t *shard_begin, int64_t *shard_count) {
  if (!c) return MG_EINVAL;
  mg_ctx *ctx = c->ctx;
  MG_REQUIRE(ctx, cfg && out_counts, "rjmcmc_array (sharded): null argument");
  const int64_t C = cfg->nchains, base = C / c->nranks, rem = C % c->nranks;
  const int64_t b = (int64_t)c->rank * base + std::min<int64_t>(c->rank, rem), cnt = base + (c->rank < rem ? 1 : 0);
  if (shard_begin) *shard_begin = b;
  if (shard_count) *shard_count = cnt;
  int64_t mine[5] = {0, 0, 0, 0, 0};          // #A, #B, accepted, cross-model proposed, cross-model accepted
  int rc = MG_OK;
  if (cnt > 0) {
    mg_rjmcmc_cfg local = *cfg;
    local.nchains = cnt; local.chain_offset = cfg->chain_offset + (uint64_t)b;
    const int64_t acc0 = ctx->naccept;
    rc = mg_rjmcmc_array(ctx, A, B, &local, a0, b0, out_model, out_samples, mine);
    mine[2] = ctx->naccept - acc0; mine[3] = ctx->rj_cross[0]; mine[4] = ctx->rj_cross[1];
  } else {
    ctx->epoch++;                              // every rank consumes the same epoch of its context's key
  }
  int64_t status = rc;
  std::vector<int64_t> all((size_t)6 * c->nranks);
  int64_t rec[6] = {mine[0], mine[1], mine[2], mine[3], mine[4], status};
  int rc2 = mg_comm_allgather(c, rec, all.data(), sizeof rec);
  if (rc) return rc;
  if (rc2) return rc2;
  int64_t tot[5] = {0, 0, 0, 0, 0};
  for (int r = 0; r < c->nranks; ++r) {
    if (all[6 * r + 5] != 0) return set_err(ctx, (int)all[6 * r + 5], "rjmcmc_array (sharded): rank %d failed", r);
    for (int k = 0; k < 5; ++k) tot[k] += all[6 * r + k];
  }
  out_counts[0] = tot[0]; out_counts[1] = tot[1];
  ctx->rj_cross[0] = tot[3]; ctx->rj_cross[1] = tot[4];
  return MG_OK;
}

// Stats.multi_mean / multi_std (stats.ml:58-87) of samples held by several ranks, from each rank's own
// (count, mean[F], std[F]): the pairs are all-gathered and pooled in rank order with the parallel-variance update
// (Chan et al.), std with n - 1 as the reference.  Every rank receives the same values.
extern "C" int mg_comm_pool_moments(mg_comm *c, int64_t n_local, const double *mean_local, const double *std_local,
                                    int32_t F, int64_t *n_total, double *out_mean, double *out_std) {
  if (!c) return MG_EINVAL;
  mg_ctx *ctx = c->ctx;
  MG_REQUIRE(ctx, mean_local && std_local && F >= 1 && F <= 4096 && n_local >= 0, "pool_moments: bad arguments");
  std::vector<double> mine((size_t)1 + 2 * F), all(((size_t)1 + 2 * F) * c->nranks);
  mine[0] = (double)n_local;
  for (int i = 0; i < F; ++i) {
    mine[1 + i] = mean_local[i];
    mine[1 + F + i] = n_local > 1 ? std_local[i] * std_local[i] * ((double)n_local - 1.0) : 0.0;   // sum of squared deviations
  }
  int rc = mg_comm_allgather(c, mine.data(), all.data(), (int64_t)(sizeof(double) * mine.size()));
  if (rc) return rc;
  double n = 0.0;
  std::vector<double> mean(F, 0.0), m2(F, 0.0);
  for (int r = 0; r < c->nranks; ++r) {
    const double *rec = all.data() + (size_t)r * mine.size();
    const double cnt = rec[0];
    if (cnt == 0.0) continue;
    const double tot = n + cnt;
    for (int i = 0; i < F; ++i) {
      const double delta = rec[1 + i] - mean[i];
      mean[i] = mean[i] + delta * (cnt / tot);
      m2[i] = m2[i] + rec[1 + F + i] + delta * delta * (n * cnt / tot);
    }
    n = tot;
  }
  if (n_total) *n_total = (int64_t)n;
  for (int i = 0; i < F; ++i) {
    if (out_mean) out_mean[i] = mean[i];
    if (out_std) out_std[i] = n > 1.0 ? sqrt(m2[i] / (n - 1.0)) : 0.0;
  }
  return MG_OK;
}
