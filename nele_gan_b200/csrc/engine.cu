// libnele_score.so -- C ABI (include/nele_score.h) and host orchestration.
//
// One engine per (process, device).  A call splits the batch into sub-batches
// ("chunks") that bound the device workspace, stages the chunk's waveforms in
// HBM, queues the three metric pipelines on one stream and copies the per-pair
// records back.  There is no CPU implementation of any metric in this library:
// without a CUDA device nele_create fails.
#include "../../include/nele_score.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "host_tables.hpp"
#include "kernels.h"

namespace nele {
void haspi_upload_tables(const float* cepm, const int* nhalf, const int* off, const float* taps, int ntaps,
                         cudaStream_t s);
}
using namespace nele;

static thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct KernelStat {
  std::string name;
  double ms = 0.0;
  int64_t launches = 0;
};

struct nele_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t s_estoi = nullptr, s_siib = nullptr;  // the three metric pipelines run concurrently
  cudaStream_t s_copy = nullptr;                       // host -> device input uploads
  cudaEvent_t ev_in[2] = {nullptr, nullptr};
  // nele_prefetch: chunk 0 of an upcoming call already uploaded (or uploading) into a staging slot
  struct Prefetched {
    bool valid = false;
    const float *ref = nullptr, *deg = nullptr;
    int n = 0;
    int64_t tot = 0;
    uint64_t seq = 0;
    uint64_t geom_hash = 0;   // FNV-1a of offs[] and lens[]: the same pointers with another layout must not match
    uint64_t born = 0;        // value of `calls` when the prefetch was issued; dropped when two calls old
    bool pcm = false;         // uploaded by nele_prefetch_pcm16
  } pf[2];
  uint64_t pf_seq = 0;
  uint64_t calls = 0;         // nele_score_batch calls with host inputs so far
  std::mutex mu;              // entry points on one engine are serialised (the workspace is shared)
  int next_slot = 0;  // staging slot the next call (or prefetch) starts with
  cudaEvent_t ev_fork = nullptr, ev_estoi = nullptr, ev_siib = nullptr;
  int concurrent = -1;                                 // NELE_CONCURRENT: 1 / 0 force the three metric pipelines onto concurrent
                                                       // streams / one stream; unset (-1): concurrent for chunks below 1024 pairs
  std::string err;
  bool f64 = false;  // recurrence precision of the ear model (NELE_HASPI_F64=1)

  // constant tables
  DevBuf bands, rs_taps, st_taps;
  int rs_fs = 0, rs_up = 1, rs_down = 1;
  int st_fs = 0, st_up = 1, st_down = 1, st_K = 0;
  double hl_cached[6] = {-1, -1, -1, -1, -1, -1};
  bool hl_ref_cached = false;

  // workspace (grow-only)
  DevBuf in_ref[2], in_deg[2], geom, sgeom, dither;  // inputs double-buffered: chunk k + 1 uploads while chunk k computes
  DevBuf in_pcm[2][3];                               // int16 staging of nele_score_batch_pcm16 (clean, enhanced, noise)
  const int16_t* pcm_noise = nullptr;                // set while a PCM-16 call runs: upload_chunk takes (clean, enhanced, noise) int16
  size_t pcm_elems[2] = {0, 0};                      // samples staged as int16 in each slot
  DevBuf x24, mid, bw, shift, envlp, rowsel, nsel, cep, cepmean, modsum;            // HASPI
  DevBuf v1_bm, v1_segsum, v1_cov, v1_msx, v1_xsum, v1_cepcorr, v1_cov3, v1_status, v1_cave, v1_ave, v1_sync5;  // HASPI version 1
  DevBuf x10, st_energy, st_kept, st_nkept, st_tob;                                 // ESTOI
  DevBuf sb_wrapdb, sb_M, sb_wact, sb_mean, sb_xdb, sb_act, sb_aidx, sb_src, sb_Fa, sb_Pact, sb_perflag, sb_lograw, sb_logspec;      // SIIB, per chunk
  DevBuf sb_base, sb_Sxx, sb_Sxy, sb_Syy, sb_Lc, sb_G, sb_perm;                     // SIIB, per sub-chunk
  DevBuf sb_rank, sb_sweeps, sb_lambda, sb_rho, sb_info;
  DevBuf kn_xk, kn_info, kn_digamma;  // SIIB k-NN estimator
  DevBuf eg_vec, eg_zt, eg_gram, eg_refl;             // SIIB tridiagonal eigen-solver: 5 x [sub][448] doubles, [sub][420][448] floats
  DevBuf out_haspi, out_raw, out_hst, out_estoi, out_est, out_siib, out_sst;
  DevBuf ft_wav, ft_geom, ft_band, ft_mag, ft_phase, ft_psd;   // feature front-end (nele_features)
  std::vector<DevBuf*> all_bufs;
  char *h_geom = nullptr, *h_sgeom = nullptr;  // pinned staging of the geometry blobs (read by blob_copy_kernel)
  size_t h_geom_cap = 0, h_sgeom_cap = 0;
  int32_t* h_M = nullptr;  // pinned
  size_t h_M_cap = 0;
  cudaEvent_t ev_M = nullptr;

  // geometry of the last chunk (for nele_get_stage)
  bool stages_valid = false;
  uint32_t stage_metrics = 0;
  std::vector<int64_t> g_off16, g_off24, g_offsub, g_off10, g_offfr, g_offW, g_offF, g_F, g_offblk;
  std::vector<int32_t> g_len16, g_n24, g_nsub, g_n10, g_nfa, g_M;
  int64_t tot24 = 0, totsub = 0, tot10 = 0, totfr = 0, totF = 0, totblk = 0;
  bool stage_v1 = false;
  int chunk_n = 0, sub_lo = 0, sub_n = 0;

  bool profiling = false;
  KernelTimer kt;
  bool kt_events = false;
  std::vector<KernelStat> kstats;
  double last_kernel_ms = 0.0;
  int64_t last_launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

static int fail(nele_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  else g_create_error = buf;
  return code;
}

#define CU(e, call)                                                                          \
  do {                                                                                       \
    cudaError_t _r = (call);                                                                 \
    if (_r != cudaSuccess)                                                                   \
      return fail(e, NELE_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

static int reserve(nele_engine* e, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return NELE_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t r = cudaMalloc(&b.p, want);
  if (r != cudaSuccess) {
    cudaGetLastError();
    return fail(e, NELE_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(r));
  }
  b.cap = want;
  return NELE_OK;
}
#define RESERVE(e, buf, bytes)                    \
  do {                                            \
    int _rc = reserve(e, buf, bytes);             \
    if (_rc != NELE_OK) return _rc;               \
  } while (0)

extern "C" int nele_abi_version(void) { return NELE_ABI_VERSION; }

extern "C" const char* nele_last_error(const nele_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

static void upload_metric_tables(cudaStream_t s) {
  {
    float cepm[kBands * kNumCep];
    host::make_cep_basis(cepm);
    host::ModFilters mf;
    host::make_mod_filters(mf);
    haspi_upload_tables(cepm, mf.nhalf, mf.offset, mf.taps.data(), (int)mf.taps.size(), s);
    haspi_v1_upload_tables(cepm, s);
  }
  {
    float win[256], tw[512];
    for (int i = 0; i < 256; ++i) win[i] = (float)(0.5 - 0.5 * cos(2.0 * host::kPi * (i + 1) / 257.0));  // hanning(258)[1:-1]
    for (int k = 0; k < 256; ++k) {
      tw[2 * k] = (float)cos(2.0 * host::kPi * k / 512.0);
      tw[2 * k + 1] = (float)(-sin(2.0 * host::kPi * k / 512.0));
    }
    int lo[15], hi[15];
    host::make_thirdoct_bins(lo, hi);
    estoi_upload_tables(win, lo, hi, tw, s);
  }
  {
    std::vector<float> win(400), decay(16), g2(host::kSiibBands * host::kSiibBins), g2t(host::kSiibBins * 32, 0.f), tw(800);
    for (int i = 0; i < 400; ++i) win[i] = (float)(0.5 - 0.5 * cos(2.0 * host::kPi * i / 400.0));  // get_window('hann', 400)
    for (int d = 0; d < 16; ++d) decay[d] = (float)(log((double)(d + 1)) / log(16.0));
    host::make_siib_g2(g2.data());
    for (int j = 0; j < host::kSiibBands; ++j)
      for (int k = 0; k < host::kSiibBins; ++k) g2t[k * 32 + j] = g2[j * host::kSiibBins + k];
    for (int k = 0; k < 400; ++k) {
      tw[2 * k] = (float)cos(2.0 * host::kPi * k / 400.0);
      tw[2 * k + 1] = (float)(-sin(2.0 * host::kPi * k / 400.0));
    }
    siib_upload_tables(win.data(), decay.data(), g2t.data(), tw.data(), s);
  }
}

extern "C" int nele_create(int device, nele_engine** out) {
  if (!out) return fail(nullptr, NELE_E_ARG, "nele_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t r = cudaGetDeviceCount(&ndev);
  if (r != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, NELE_E_NODEVICE,
                "nele_create: no CUDA device (%s); libnele_score has no CPU path",
                r != cudaSuccess ? cudaGetErrorString(r) : "device count is 0");
  }
  if (device < 0 || device >= ndev) return fail(nullptr, NELE_E_ARG, "nele_create: device %d out of range [0,%d)", device, ndev);
  nele_engine* e = new nele_engine();
  e->device = device;
  const char* p = getenv("NELE_HASPI_F64");
  e->f64 = (p && p[0] == '1');
  // One stream for large chunks: every kernel of the three pipelines fills the SMs on its own (register-
  // or shared-memory-limited occupancy), so concurrent streams gain nothing there (302 vs 307 ms per 4096
  // full-rank pairs) and blur the per-kernel timings.  Chunks below 1024 pairs fork / join the three metric
  // pipelines (see score_core); NELE_CONCURRENT=0 / 1 forces either behaviour.
  p = getenv("NELE_CONCURRENT");
  e->concurrent = (p && (p[0] == '0' || p[0] == '1')) ? p[0] - '0' : -1;
  e->all_bufs = {&e->bands, &e->rs_taps, &e->st_taps, &e->in_ref[0], &e->in_ref[1], &e->in_deg[0], &e->in_deg[1], &e->geom, &e->sgeom, &e->dither,
                 &e->in_pcm[0][0], &e->in_pcm[0][1], &e->in_pcm[0][2], &e->in_pcm[1][0], &e->in_pcm[1][1], &e->in_pcm[1][2],
                 &e->x24, &e->mid, &e->bw, &e->shift, &e->envlp, &e->rowsel, &e->nsel, &e->cep, &e->cepmean, &e->modsum,
                 &e->v1_bm, &e->v1_segsum, &e->v1_cov, &e->v1_msx, &e->v1_xsum, &e->v1_cepcorr, &e->v1_cov3, &e->v1_status, &e->v1_cave, &e->v1_ave, &e->v1_sync5,
                 &e->x10, &e->st_energy, &e->st_kept, &e->st_nkept, &e->st_tob,
                 &e->sb_wrapdb, &e->sb_M, &e->sb_wact, &e->sb_mean, &e->sb_xdb, &e->sb_act, &e->sb_aidx, &e->sb_src, &e->sb_Fa, &e->sb_Pact, &e->sb_perflag, &e->sb_lograw, &e->sb_logspec,
                 &e->sb_base, &e->sb_Sxx, &e->sb_Sxy, &e->sb_Syy, &e->sb_Lc, &e->sb_G, &e->sb_perm,
                 &e->sb_rank, &e->sb_sweeps, &e->sb_lambda, &e->sb_rho, &e->sb_info, &e->kn_xk, &e->kn_info, &e->kn_digamma, &e->eg_vec, &e->eg_zt, &e->eg_gram, &e->eg_refl,
                 &e->out_haspi, &e->out_raw, &e->out_hst, &e->out_estoi, &e->out_est, &e->out_siib, &e->out_sst,
                 &e->ft_wav, &e->ft_geom, &e->ft_band, &e->ft_mag, &e->ft_phase, &e->ft_psd};
#define CUC(call)                                                                             \
  do {                                                                                        \
    cudaError_t _r = (call);                                                                  \
    if (_r != cudaSuccess) {                                                                  \
      fail(nullptr, NELE_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(_r));             \
      delete e;                                                                               \
      return NELE_E_CUDA;                                                                     \
    }                                                                                         \
  } while (0)
  CUC(cudaSetDevice(device));
  CUC(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CUC(cudaEventCreate(&e->ev0));
  CUC(cudaEventCreate(&e->ev1));
  CUC(cudaEventCreateWithFlags(&e->ev_M, cudaEventDisableTiming));
  CUC(cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
  CUC(cudaEventCreateWithFlags(&e->ev_in[0], cudaEventDisableTiming));
  CUC(cudaEventCreateWithFlags(&e->ev_in[1], cudaEventDisableTiming));
  CUC(cudaStreamCreateWithFlags(&e->s_estoi, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&e->s_siib, cudaStreamNonBlocking));
  CUC(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CUC(cudaEventCreateWithFlags(&e->ev_estoi, cudaEventDisableTiming));
  CUC(cudaEventCreateWithFlags(&e->ev_siib, cudaEventDisableTiming));
  haspi_upload_constants(e->stream);
  upload_metric_tables(e->stream);
  CUC(cudaGetLastError());
#undef CUC
  *out = e;
  return NELE_OK;
}

extern "C" void nele_destroy(nele_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (DevBuf* b : e->all_bufs)
    if (b->p) cudaFree(b->p);
  if (e->h_geom) cudaFreeHost(e->h_geom);
  if (e->h_sgeom) cudaFreeHost(e->h_sgeom);
  if (e->h_M) cudaFreeHost(e->h_M);
  if (e->kt_events)
    for (int i = 0; i < KernelTimer::kMax; ++i) {
      cudaEventDestroy(e->kt.ev0[i]);
      cudaEventDestroy(e->kt.ev1[i]);
    }
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->ev_M) cudaEventDestroy(e->ev_M);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_estoi) cudaEventDestroy(e->ev_estoi);
  if (e->ev_siib) cudaEventDestroy(e->ev_siib);
  if (e->s_copy) cudaStreamDestroy(e->s_copy);
  if (e->ev_in[0]) cudaEventDestroy(e->ev_in[0]);
  if (e->ev_in[1]) cudaEventDestroy(e->ev_in[1]);
  if (e->s_estoi) cudaStreamDestroy(e->s_estoi);
  if (e->s_siib) cudaStreamDestroy(e->s_siib);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

extern "C" int nele_set_profiling(nele_engine* e, int on) {
  if (!e) return NELE_E_ARG;
  std::lock_guard<std::mutex> lock(e->mu);
  CU(e, cudaSetDevice(e->device));
  if (on && !e->kt_events) {
    for (int i = 0; i < KernelTimer::kMax; ++i) {
      CU(e, cudaEventCreate(&e->kt.ev0[i]));
      CU(e, cudaEventCreate(&e->kt.ev1[i]));
    }
    e->kt_events = true;
  }
  e->profiling = on != 0;
  return NELE_OK;
}

static int ensure_tables(nele_engine* e, int fs, bool haspi_rate_ok, const double* hl, bool ref_gets_hl, cudaStream_t s) {
  double h[6] = {0, 0, 0, 0, 0, 0};
  if (hl) memcpy(h, hl, sizeof(h));
  if (!e->bands.p || memcmp(h, e->hl_cached, sizeof(h)) != 0 || e->hl_ref_cached != ref_gets_hl) {
    BandConst bc[kBands];
    host::make_band_consts(h, bc);
    if (ref_gets_hl)  // hasqi_v2 calls eb_EarModel with itype = 2: the reference signal gets the audiogram too (pyhaspi2.py:1162-1165)
      for (int k = 0; k < kBands; ++k) {
        bc[k].attn_ohc[0] = bc[k].attn_ohc[1];
        bc[k].bwmin[0] = bc[k].bwmin[1];
        bc[k].lowknee[0] = bc[k].lowknee[1];
        bc[k].cr[0] = bc[k].cr[1];
        bc[k].attn_ihc[0] = bc[k].attn_ihc[1];
      }
    e->hl_ref_cached = ref_gets_hl;
    RESERVE(e, e->bands, sizeof(bc));
    CU(e, cudaMemcpyAsync(e->bands.p, bc, sizeof(bc), cudaMemcpyHostToDevice, s));
    CU(e, cudaStreamSynchronize(s));
    memcpy(e->hl_cached, h, sizeof(h));
  }
  const int hfs = haspi_rate_ok ? fs : kFs24;
  if (e->rs_fs != hfs) {
    host::ResampyTaps rt;
    if (hfs == kFs24) {
      rt.up = rt.down = 1;
      rt.taps.assign(128, 0.0);
    } else {
      host::make_resampy_taps(hfs, kFs24, rt);
    }
    RESERVE(e, e->rs_taps, rt.taps.size() * sizeof(double));
    CU(e, cudaMemcpyAsync(e->rs_taps.p, rt.taps.data(), rt.taps.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(e, cudaStreamSynchronize(s));
    e->rs_fs = hfs;
    e->rs_up = rt.up;
    e->rs_down = rt.down;
  }
  if (e->st_fs != fs) {
    host::PolyTaps pt;
    if (fs == 10000) {
      pt.up = pt.down = 1;
      pt.K = 0;
      pt.taps.assign(1, 1.0);
    } else {
      host::make_estoi_polytaps(fs, 10000, pt);
    }
    RESERVE(e, e->st_taps, pt.taps.size() * sizeof(double));
    CU(e, cudaMemcpyAsync(e->st_taps.p, pt.taps.data(), pt.taps.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(e, cudaStreamSynchronize(s));
    estoi_upload_polytaps(pt.taps.data(), pt.up, pt.K, s);
    e->st_fs = fs;
    e->st_up = pt.up;
    e->st_down = pt.down;
    e->st_K = pt.K;
  }
  return NELE_OK;
}

static const double kNaN = nan("");

// pack per-pair host arrays into one blob -> one H2D copy
struct GeomPacker {
  std::vector<char> blob;
  size_t add(const void* src, size_t bytes) {
    const size_t off = (blob.size() + 15) & ~(size_t)15;
    blob.resize(off + bytes);
    memcpy(blob.data() + off, src, bytes);
    return off;
  }
};

// The per-chunk geometry arrays reach the device through a copy *kernel* that reads pinned host
// memory: a cudaMemcpyAsync would queue on the host-to-device copy engine behind whatever
// waveform upload (next chunk, nele_prefetch) is in flight there and stall the kernels for tens
// of milliseconds.
__global__ void blob_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
static int push_blob(nele_engine* e, const std::vector<char>& blob, char** pinned, size_t* cap, DevBuf& dev, cudaStream_t s) {
  const size_t bytes = (blob.size() + 15) & ~(size_t)15;
  if (*cap < bytes) {
    if (*pinned) cudaFreeHost(*pinned);
    *pinned = nullptr;
    *cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    CU(e, cudaMallocHost((void**)pinned, want));
    *cap = want;
  }
  memcpy(*pinned, blob.data(), blob.size());
  RESERVE(e, dev, bytes);
  const size_t n16 = bytes / 16;
  blob_copy_kernel<<<(unsigned)std::min<size_t>((n16 + 255) / 256, 64), 256, 0, s>>>((const uint4*)*pinned, (uint4*)dev.p, n16);
  CU(e, cudaGetLastError());
  return NELE_OK;
}

static void collect_kernel_times(nele_engine* e) {
  // called after the stream has been synchronised
  for (int i = 0; i < e->kt.count; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->kt.ev0[i], e->kt.ev1[i]) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    KernelStat* st = nullptr;
    for (auto& k : e->kstats)
      if (k.name == e->kt.names[i]) st = &k;
    if (!st) {
      e->kstats.push_back(KernelStat());
      st = &e->kstats.back();
      st->name = e->kt.names[i];
    }
    st->ms += ms;
    st->launches += 1;
  }
  e->kt.count = 0;
}

// One sub-batch of a call: pairs [first, last), where its waveforms sit in the caller's buffers
// and where they go in the staging buffers.
struct ChunkPlan {
  int first = 0, last = 0;
  int64_t tot = 0, lo = 0, hi = 0, packed = 0;
  bool span_copy = false;
  std::vector<int64_t> off16;  // start of every pair in the staged (or the caller's device) buffers
};

static void plan_chunks(const int64_t* offs, const int32_t* lens, int n, bool dev_in, int max_pairs, int64_t max_samples,
                        std::vector<ChunkPlan>& plans) {
  int first = 0;
  while (first < n) {
    ChunkPlan c;
    c.first = first;
    int last = first;
    c.lo = offs[first];
    c.hi = offs[first] + lens[first];
    while (last < n && last - first < max_pairs && (last == first || c.tot + lens[last] <= max_samples)) {
      c.tot += lens[last];
      c.lo = std::min(c.lo, offs[last]);
      c.hi = std::max(c.hi, offs[last] + (int64_t)lens[last]);
      ++last;
    }
    c.last = last;
    c.span_copy = !dev_in && (c.hi - c.lo) <= 2 * c.tot + 4096;
    c.off16.resize(last - first);
    for (int i = first; i < last; ++i) {
      c.off16[i - first] = dev_in ? offs[i] : (c.span_copy ? offs[i] - c.lo : c.packed);
      c.packed += (lens[i] + 3) & ~3;
    }
    plans.push_back(std::move(c));
    first = last;
  }
}

// PCM-16 inputs (what the reference's WAV files hold): ref = clean / 32768 (librosa.load, audio_util.py:128-130),
// deg = enhanced / 32768 + noise / 32768 (audio_util.py:139) -- both exact in float32
__global__ void pcm16_to_float_kernel(const int16_t* __restrict__ c, const int16_t* __restrict__ en, const int16_t* __restrict__ nz,
                                      float* __restrict__ ref, float* __restrict__ deg, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    ref[i] = (float)c[i] * (1.f / 32768.f);
    deg[i] = (float)en[i] * (1.f / 32768.f) + (float)nz[i] * (1.f / 32768.f);
  }
}

// queue the host -> device upload of one chunk's waveforms on the copy stream
static int upload_chunk(nele_engine* e, const ChunkPlan& c, int slot, const float* ref, const float* deg,
                        const int64_t* offs, const int32_t* lens) {
  const size_t in_elems = c.span_copy ? (size_t)(c.hi - c.lo) : (size_t)c.packed;
  RESERVE(e, e->in_ref[slot], in_elems * sizeof(float));
  RESERVE(e, e->in_deg[slot], in_elems * sizeof(float));
  cudaStream_t sc = e->s_copy;
  if (e->pcm_noise) {
    // int16 triple: `ref` / `deg` carry the clean / enhanced int16 pointers; 6 instead of 8 bytes per sample cross PCIe
    const int16_t* src[3] = {reinterpret_cast<const int16_t*>(ref), reinterpret_cast<const int16_t*>(deg), e->pcm_noise};
    for (int k = 0; k < 3; ++k) {
      RESERVE(e, e->in_pcm[slot][k], in_elems * sizeof(int16_t));
      if (c.span_copy) {
        CU(e, cudaMemcpyAsync(e->in_pcm[slot][k].p, src[k] + c.lo, in_elems * sizeof(int16_t), cudaMemcpyHostToDevice, sc));
      } else {
        for (int i = c.first; i < c.last; ++i)
          CU(e, cudaMemcpyAsync((int16_t*)e->in_pcm[slot][k].p + c.off16[i - c.first], src[k] + offs[i], sizeof(int16_t) * lens[i],
                                cudaMemcpyHostToDevice, sc));
      }
    }
    // the expansion to float32 runs on the compute stream at the start of the chunk (score_core): launched here, on the
    // copy stream, it queued behind whole compute kernels and delayed the next step by ~5 ms
    e->pcm_elems[slot] = in_elems;
    CU(e, cudaEventRecord(e->ev_in[slot], sc));
    return NELE_OK;
  }
  if (c.span_copy) {
    CU(e, cudaMemcpyAsync(e->in_ref[slot].p, ref + c.lo, in_elems * sizeof(float), cudaMemcpyHostToDevice, sc));
    CU(e, cudaMemcpyAsync(e->in_deg[slot].p, deg + c.lo, in_elems * sizeof(float), cudaMemcpyHostToDevice, sc));
  } else {
    for (int i = c.first; i < c.last; ++i) {
      CU(e, cudaMemcpyAsync((float*)e->in_ref[slot].p + c.off16[i - c.first], ref + offs[i], sizeof(float) * lens[i], cudaMemcpyHostToDevice, sc));
      CU(e, cudaMemcpyAsync((float*)e->in_deg[slot].p + c.off16[i - c.first], deg + offs[i], sizeof(float) * lens[i], cudaMemcpyHostToDevice, sc));
    }
  }
  CU(e, cudaEventRecord(e->ev_in[slot], sc));
  return NELE_OK;
}

static uint64_t geom_hash(const int64_t* offs, const int32_t* lens, int n) {
  uint64_t h = 1469598103934665603ULL;
  auto mix = [&h](uint64_t v) {
    for (int k = 0; k < 8; ++k) {
      h ^= (v >> (8 * k)) & 0xff;
      h *= 1099511628211ULL;
    }
  };
  for (int i = 0; i < n; ++i) {
    mix((uint64_t)offs[i]);
    mix((uint64_t)(uint32_t)lens[i]);
  }
  return h;
}

static int score_core(nele_engine* e, const float* ref, const float* deg, const int64_t* offs, const int32_t* lens, int n,
                      int fs, uint32_t metrics, uint32_t flags, const float* dither, int64_t dither_rows, uint64_t seed,
                      const double* hl, double* scores, double* haspi_raw, int32_t* status, void* stream);

extern "C" int nele_score_batch(nele_engine* e, const float* ref, const float* deg, const int64_t* offs,
                                const int32_t* lens, int n, int fs, uint32_t metrics, uint32_t flags,
                                const float* dither, int64_t dither_rows, uint64_t seed, const double* hl,
                                double* scores, double* haspi_raw, int32_t* status, void* stream) {
  if (!e) return NELE_E_ARG;
  std::lock_guard<std::mutex> lock(e->mu);
  e->pcm_noise = nullptr;
  return score_core(e, ref, deg, offs, lens, n, fs, metrics, flags, dither, dither_rows, seed, hl, scores, haspi_raw, status, stream);
}

extern "C" int nele_score_batch_pcm16(nele_engine* e, const int16_t* clean, const int16_t* enhanced, const int16_t* noise,
                                      const int64_t* offs, const int32_t* lens, int n, int fs, uint32_t metrics, uint32_t flags,
                                      const float* dither, int64_t dither_rows, uint64_t seed, const double* hl,
                                      double* scores, double* haspi_raw, int32_t* status, void* stream) {
  if (!e) return NELE_E_ARG;
  std::lock_guard<std::mutex> lock(e->mu);
  if (flags & NELE_FLAG_DEVICE_INPUT) return fail(e, NELE_E_ARG, "nele_score_batch_pcm16 takes host buffers");
  if (n > 0 && !noise) return fail(e, NELE_E_ARG, "nele_score_batch_pcm16: null pointer");
  e->pcm_noise = noise;
  const int rc = score_core(e, reinterpret_cast<const float*>(clean), reinterpret_cast<const float*>(enhanced), offs, lens, n, fs,
                            metrics, flags, dither, dither_rows, seed, hl, scores, haspi_raw, status, stream);
  e->pcm_noise = nullptr;
  return rc;
}

static int score_core(nele_engine* e, const float* ref, const float* deg, const int64_t* offs, const int32_t* lens, int n,
                      int fs, uint32_t metrics, uint32_t flags, const float* dither, int64_t dither_rows, uint64_t seed,
                      const double* hl, double* scores, double* haspi_raw, int32_t* status, void* stream) {
  if (n < 0 || (n > 0 && (!ref || !deg || !offs || !lens || !scores)))
    return fail(e, NELE_E_ARG, "nele_score_batch: null pointer or negative n");
  if ((metrics & ~NELE_METRIC_ALL) || metrics == 0) return fail(e, NELE_E_ARG, "nele_score_batch: bad metric mask 0x%x", metrics);
  if (fs <= 0) return fail(e, NELE_E_ARG, "nele_score_batch: fs = %d", fs);
  if (dither && dither_rows <= 0) return fail(e, NELE_E_ARG, "nele_score_batch: dither given with dither_rows = %lld", (long long)dither_rows);
  for (int i = 0; i < n; ++i)
    if (lens[i] <= 0 || offs[i] < 0) return fail(e, NELE_E_ARG, "nele_score_batch: pair %d has length %d / offset %lld", i, lens[i], (long long)offs[i]);
  e->last_kernel_ms = 0.0;
  e->last_launches = 0;
  e->stages_valid = false;
  e->kstats.clear();
  e->kt.count = 0;
  e->kt.enabled = e->profiling && e->kt_events;
  KernelTimer* kt = e->kt.enabled ? &e->kt : nullptr;
  if (n == 0) return NELE_OK;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  const bool do_haspi = metrics & NELE_METRIC_HASPI, do_siib = metrics & NELE_METRIC_SIIB, do_estoi = metrics & NELE_METRIC_ESTOI;
  const bool haspi_rate_ok = fs <= kFs24;
  const bool siib_rate_ok = fs == 16000;  // audio_util.py:131,159,187 assert it; the wrapper's R = fs / 200 presumes it
  const bool dev_in = flags & NELE_FLAG_DEVICE_INPUT;
  const bool mapped = flags & NELE_FLAG_MAPPED;
  const bool hasqi = do_haspi && (flags & NELE_FLAG_HASQI_V2);
  const bool haspi_v1 = do_haspi && ((flags & NELE_FLAG_HASPI_V1) || hasqi);  // HASQI runs on the version-1 pipeline
  int rc = ensure_tables(e, fs, haspi_rate_ok, hl, hasqi, s);
  if (rc != NELE_OK) return rc;

  if (dither && !(flags & NELE_FLAG_NO_DITHER)) {
    const size_t bytes = (size_t)2 * dither_rows * kBands * sizeof(float);
    RESERVE(e, e->dither, bytes);
    CU(e, cudaMemcpyAsync(e->dither.p, dither, bytes, cudaMemcpyHostToDevice, s));
  }
  for (int i = 0; i < n; ++i) {
    scores[3 * i + 0] = scores[3 * i + 1] = scores[3 * i + 2] = kNaN;
    if (status) status[i] = (NELE_ST_SKIPPED) | (NELE_ST_SKIPPED << 8) | (NELE_ST_SKIPPED << 16);
  }

  // ---- chunking: bound the workspace by pairs and by total samples.  With host inputs the upload
  // of chunk k + 1 (copy engine, own stream, second staging buffer) hides behind the kernels of
  // chunk k.  (Cutting a 4096-pair call into two 2048-pair chunks to hide half of its own upload
  // was measured and lost: 348.8 vs 340.9 ms end to end -- the kernels want the larger batch.)
  const int64_t kMaxChunkSamples = 256LL * 1000 * 1000;  // input-rate samples per signal
  // HASPI version 1 stages the basilar-membrane motion of both signals in HBM (384 bytes per input
  // sample at 16 kHz): smaller chunks
  const int64_t kMaxChunkSamplesV1 = 64LL * 1000 * 1000;
  const int kMaxChunkPairs = haspi_v1 ? 2048 : 4096;
  int kHostChunkPairs = kMaxChunkPairs;
  if (const char* hc = getenv("NELE_HOST_CHUNK")) {  // tuning knob: pairs per chunk when the inputs come from the host
    const int v = atoi(hc);
    if (v > 0) kHostChunkPairs = std::min(v, kMaxChunkPairs);
  }
  const int kSiibSub = 4096;                              // pairs per SIIB matrix sub-chunk (6.2 MB of matrices per pair)
  static const bool trace = [] { const char* p = getenv("NELE_TRACE"); return p && p[0] == '1'; }();
  const auto t_call = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };
  std::vector<ChunkPlan> plans;
  plan_chunks(offs, lens, n, dev_in, dev_in ? kMaxChunkPairs : kHostChunkPairs,
              haspi_v1 ? kMaxChunkSamplesV1 : kMaxChunkSamples, plans);
  int slot0 = e->next_slot;
  if (!dev_in) {
    // chunk 0 may already be on its way (nele_prefetch): take the oldest matching slot.  A prefetch matches by
    // buffer addresses *and* layout (hash of offs / lens), and expires when two calls have gone by without
    // consuming it (a skipped step, an exception in the caller): freed-and-reallocated host buffers at the same
    // address then cannot pick up stale waveforms.  nele_prefetch_cancel drops pending prefetches explicitly.
    ++e->calls;
    const uint64_t gh = geom_hash(offs, lens, n);
    for (int k = 0; k < 2; ++k)
      if (e->pf[k].valid && e->calls - e->pf[k].born > 2) e->pf[k].valid = false;
    int hit = -1;
    for (int k = 0; k < 2; ++k)
      if (e->pf[k].valid && e->pf[k].ref == ref && e->pf[k].deg == deg && e->pf[k].n == n && e->pf[k].tot == plans[0].tot &&
          e->pf[k].geom_hash == gh && e->pf[k].pcm == (e->pcm_noise != nullptr) && (hit < 0 || e->pf[k].seq < e->pf[hit].seq))
        hit = k;
    if (hit >= 0) {
      slot0 = hit;
      e->pf[hit].valid = false;
      if (plans.size() > 1) e->pf[hit ^ 1].valid = false;  // the second chunk will overwrite the other slot
    } else {
      if (e->pf[slot0].valid) slot0 ^= 1;                   // do not clobber a pending prefetch
      if (e->pf[slot0].valid || plans.size() > 1) e->pf[0].valid = e->pf[1].valid = false;
      rc = upload_chunk(e, plans[0], slot0, ref, deg, offs, lens);
      if (rc != NELE_OK) return rc;
    }
    if (trace) fprintf(stderr, "[nele] first upload %s at %.2f ms\n", hit >= 0 ? "was prefetched" : "queued", since());
    e->next_slot = (slot0 + (int)plans.size()) & 1;
  }
  for (size_t ci = 0; ci < plans.size(); ++ci) {
    const ChunkPlan& cp = plans[ci];
    const int first = cp.first, last = cp.last;
    const int slot = (slot0 + (int)ci) & 1;
    const int cn = last - first;
    // ---- geometry
    e->g_off16.resize(cn); e->g_len16.resize(cn); e->g_off24.resize(cn); e->g_n24.resize(cn);
    e->g_offsub.resize(cn); e->g_nsub.resize(cn); e->g_off10.resize(cn); e->g_n10.resize(cn);
    e->g_offfr.resize(cn); e->g_nfa.resize(cn); e->g_offW.resize(cn); e->g_offblk.resize(cn);
    int64_t t24 = 0, tsub = 0, t10 = 0, tfr = 0, tW = 0, tblk = 0;
    int max_nsub = 0, max_n10 = 0, max_nfa = 0, max_n24 = 0;
    for (int i = 0; i < cn; ++i) {
      const int L = lens[first + i];
      e->g_len16[i] = L;
      e->g_off16[i] = cp.off16[i];
      const int n24 = (fs == kFs24 || !haspi_rate_ok) ? L : (int)(((int64_t)L * e->rs_up + e->rs_down - 1) / e->rs_down);
      const int nsub = (n24 + kDecim - 1) / kDecim;
      e->g_n24[i] = n24;
      e->g_nsub[i] = nsub;
      e->g_off24[i] = t24;
      e->g_offsub[i] = tsub;
      e->g_offblk[i] = tblk;
      t24 += (n24 + 31) & ~31;
      tsub += nsub;
      tblk += (n24 + 191) / 192;
      max_nsub = std::max(max_nsub, nsub);
      max_n24 = std::max(max_n24, n24);
      const int n10 = (int)(((int64_t)L * e->st_up + e->st_down - 1) / e->st_down);
      const int nfa = n10 > 256 ? (n10 - 256 + 127) / 128 : 0;
      e->g_n10[i] = n10;
      e->g_nfa[i] = nfa;
      e->g_off10[i] = t10;
      e->g_offfr[i] = tfr;
      t10 += (n10 + 31) & ~31;
      tfr += nfa;
      max_n10 = std::max(max_n10, n10);
      max_nfa = std::max(max_nfa, nfa);
      const int Lp = std::max(L, 401);
      e->g_offW[i] = tW;
      tW += (Lp - 400 + 199) / 200;
    }
    e->tot24 = t24;
    e->totsub = tsub;
    e->tot10 = t10;
    e->totfr = tfr;
    e->totblk = tblk;
    e->chunk_n = cn;

    // ---- inputs
    const float *d_ref = ref, *d_deg = deg;
    if (!dev_in) {
      CU(e, cudaStreamWaitEvent(s, e->ev_in[slot], 0));
      d_ref = (const float*)e->in_ref[slot].p;
      d_deg = (const float*)e->in_deg[slot].p;
      if (e->pcm_noise) {
        pcm16_to_float_kernel<<<1184, 256, 0, s>>>((const int16_t*)e->in_pcm[slot][0].p, (const int16_t*)e->in_pcm[slot][1].p,
                                                   (const int16_t*)e->in_pcm[slot][2].p, (float*)e->in_ref[slot].p,
                                                   (float*)e->in_deg[slot].p, e->pcm_elems[slot]);
        CU(e, cudaGetLastError());
        ++e->last_launches;
      }
    }
    // ---- geometry arrays -> device (one blob, one copy)
    GeomPacker gp;
    const size_t o_off16 = gp.add(e->g_off16.data(), cn * sizeof(int64_t));
    const size_t o_off24 = gp.add(e->g_off24.data(), cn * sizeof(int64_t));
    const size_t o_offsub = gp.add(e->g_offsub.data(), cn * sizeof(int64_t));
    const size_t o_off10 = gp.add(e->g_off10.data(), cn * sizeof(int64_t));
    const size_t o_offfr = gp.add(e->g_offfr.data(), cn * sizeof(int64_t));
    const size_t o_offW = gp.add(e->g_offW.data(), cn * sizeof(int64_t));
    const size_t o_offblk = gp.add(e->g_offblk.data(), cn * sizeof(int64_t));
    const size_t o_len16 = gp.add(e->g_len16.data(), cn * sizeof(int32_t));
    const size_t o_n24 = gp.add(e->g_n24.data(), cn * sizeof(int32_t));
    const size_t o_nsub = gp.add(e->g_nsub.data(), cn * sizeof(int32_t));
    const size_t o_n10 = gp.add(e->g_n10.data(), cn * sizeof(int32_t));
    const size_t o_nfa = gp.add(e->g_nfa.data(), cn * sizeof(int32_t));
    rc = push_blob(e, gp.blob, &e->h_geom, &e->h_geom_cap, e->geom, s);
    if (rc != NELE_OK) return rc;
    if (!dev_in && ci + 1 < plans.size()) {  // the other slot was released when chunk ci - 1 finished
      rc = upload_chunk(e, plans[ci + 1], slot ^ 1, ref, deg, offs, lens);
      if (rc != NELE_OK) return rc;
      if (trace) fprintf(stderr, "[nele] chunk %zu: next upload queued at %.2f ms\n", ci, since());
    }
    const char* gb = (const char*)e->geom.p;
    PairGeom g;
    g.off16 = (const int64_t*)(gb + o_off16);
    g.off24 = (const int64_t*)(gb + o_off24);
    g.offsub = (const int64_t*)(gb + o_offsub);
    g.len16 = (const int32_t*)(gb + o_len16);
    g.n24 = (const int32_t*)(gb + o_n24);
    g.nsub = (const int32_t*)(gb + o_nsub);
    g.offblk = (const int64_t*)(gb + o_offblk);
    EstoiGeom eg;
    eg.off16 = g.off16;
    eg.len16 = g.len16;
    eg.off10 = (const int64_t*)(gb + o_off10);
    eg.n10 = (const int32_t*)(gb + o_n10);
    eg.offfr = (const int64_t*)(gb + o_offfr);
    eg.nfa = (const int32_t*)(gb + o_nfa);
    SiibGeom sg;
    sg.off16 = g.off16;
    sg.len16 = g.len16;
    sg.offW = (const int64_t*)(gb + o_offW);
    sg.offF = nullptr;
    sg.F = nullptr;

    const bool run_haspi = do_haspi && haspi_rate_ok, run_siib = do_siib && siib_rate_ok, run_estoi = do_estoi;
    RESERVE(e, e->out_haspi, cn * sizeof(double));
    RESERVE(e, e->out_raw, (size_t)cn * kNumMod * sizeof(double));
    RESERVE(e, e->out_hst, (size_t)cn * sizeof(int32_t));
    RESERVE(e, e->out_estoi, cn * sizeof(double));
    RESERVE(e, e->out_est, (size_t)cn * sizeof(int32_t));
    RESERVE(e, e->out_siib, cn * sizeof(double));
    RESERVE(e, e->out_sst, (size_t)cn * sizeof(int32_t));

    if (trace) fprintf(stderr, "[nele] chunk %zu: geometry on device at %.2f ms\n", ci, since());
    CU(e, cudaEventRecord(e->ev0, s));
    // fork: HASPI stays on the launching stream, ESTOI and SIIB get their own so that the
    // latency-bound kernels of one metric fill the SMs the others leave idle
    // Large chunks fill the GPU kernel by kernel and gain nothing from concurrency (measured: 302 vs 307 ms per 4096
    // full-rank pairs); small ones -- a GAN sampling round spread over eight GPUs -- are latency bound per kernel and
    // overlap well (135 pairs: 20.4 -> 17.0 ms).  Profiling (per-kernel events) keeps one stream.
    const bool serial = kt ? true : (e->concurrent >= 0 ? e->concurrent == 0 : cn >= 1024);
    cudaStream_t se = serial ? s : e->s_estoi, ss = serial ? s : e->s_siib;
    if (!serial) {
      CU(e, cudaEventRecord(e->ev_fork, s));
      CU(e, cudaStreamWaitEvent(se, e->ev_fork, 0));
      CU(e, cudaStreamWaitEvent(ss, e->ev_fork, 0));
    }
    // ---- SIIB stage 0: wrapper VAD -> tiling factors (the host needs them to size the rest)
    SiibBuffers sb;
    memset(&sb, 0, sizeof(sb));
    if (run_siib) {
      RESERVE(e, e->sb_wrapdb, (size_t)tW * sizeof(double));
      RESERVE(e, e->sb_M, (size_t)cn * sizeof(int32_t));
      RESERVE(e, e->sb_wact, (size_t)cn * sizeof(int32_t));
      if (e->h_M_cap < (size_t)cn) {
        if (e->h_M) cudaFreeHost(e->h_M);
        e->h_M = nullptr;
        CU(e, cudaMallocHost((void**)&e->h_M, (size_t)cn * sizeof(int32_t)));
        e->h_M_cap = cn;
      }
      sb.ref = d_ref;
      sb.deg = d_deg;
      sb.wrapdb = (double*)e->sb_wrapdb.p;
      sb.M = (int32_t*)e->sb_M.p;
      sb.wrap_active = (int32_t*)e->sb_wact.p;
      e->last_launches += siib_run_wrapvad(sg, sb, cn, flags & NELE_FLAG_SIIB_NO_TILE, kt, ss);
      CU(e, cudaMemcpyAsync(e->h_M, sb.M, (size_t)cn * sizeof(int32_t), cudaMemcpyDeviceToHost, ss));
      CU(e, cudaEventRecord(e->ev_M, ss));
    }
    // ---- HASPI
    HaspiBuffers hb;
    memset(&hb, 0, sizeof(hb));
    if (run_haspi) {
      RESERVE(e, e->x24, (size_t)2 * t24 * sizeof(float));
      RESERVE(e, e->mid, (size_t)2 * t24 * sizeof(float));
      RESERVE(e, e->bw, (size_t)cn * 2 * kBands * sizeof(double));
      RESERVE(e, e->shift, (size_t)cn * kBands * sizeof(int32_t));
      if (!haspi_v1) {
        RESERVE(e, e->envlp, (size_t)2 * tsub * kBands * sizeof(float));
        RESERVE(e, e->rowsel, (size_t)tsub * sizeof(int32_t));
        RESERVE(e, e->nsel, (size_t)cn * sizeof(int32_t));
        RESERVE(e, e->cep, (size_t)2 * kNumCep * tsub * sizeof(float));
        RESERVE(e, e->cepmean, (size_t)cn * 2 * kNumCep * sizeof(double));
        RESERVE(e, e->modsum, (size_t)cn * kNumCep * kNumMod * 5 * sizeof(double));
      } else {
        RESERVE(e, e->v1_bm, (size_t)2 * t24 * kBands * sizeof(float));
        RESERVE(e, e->v1_segsum, (size_t)4 * tblk * kBands * sizeof(float));
        RESERVE(e, e->v1_cov, (size_t)tblk * kBands * sizeof(float));
        RESERVE(e, e->v1_msx, (size_t)tblk * kBands * sizeof(float));
        RESERVE(e, e->v1_xsum, (size_t)tblk * sizeof(double));
        RESERVE(e, e->v1_cepcorr, (size_t)cn * sizeof(double));
        RESERVE(e, e->v1_cov3, (size_t)cn * 3 * sizeof(double));
        RESERVE(e, e->v1_status, (size_t)cn * sizeof(int32_t));
        RESERVE(e, e->v1_cave, (size_t)cn * 2 * kBands * sizeof(double));
        RESERVE(e, e->v1_ave, (size_t)cn * 2 * kBands * sizeof(double));
        RESERVE(e, e->v1_sync5, (size_t)cn * sizeof(double));
      }
      hb.ref = d_ref;
      hb.deg = d_deg;
      hb.x24 = (float*)e->x24.p;
      hb.mid = (float*)e->mid.p;
      hb.tot24 = t24;
      hb.bw = (double*)e->bw.p;
      hb.shift = (int32_t*)e->shift.p;
      hb.envlp = (float*)e->envlp.p;
      hb.totsub = tsub;
      hb.rowsel = (int32_t*)e->rowsel.p;
      hb.nsel = (int32_t*)e->nsel.p;
      hb.cep = (float*)e->cep.p;
      hb.cepmean = (double*)e->cepmean.p;
      hb.modsum = (double*)e->modsum.p;
      hb.bands = (const BandConst*)e->bands.p;
      hb.rs_taps = (const double*)e->rs_taps.p;
      hb.rs_up = e->rs_up;
      hb.rs_down = e->rs_down;
      hb.dither = (dither && !(flags & NELE_FLAG_NO_DITHER)) ? (const float*)e->dither.p : nullptr;
      hb.dither_rows = dither_rows;
      hb.seed = seed;
      hb.no_dither = (flags & NELE_FLAG_NO_DITHER) ? 1 : 0;
      hb.pair_base = first;
      if (!haspi_v1) {
        e->last_launches += haspi_run(g, hb, cn, max_nsub, e->f64, kt, s);
        e->last_launches += haspi_finish(hb, cn, (double*)e->out_haspi.p, (double*)e->out_raw.p, (int32_t*)e->out_hst.p, kt, s);
      } else {
        HaspiV1Buffers vb;
        vb.bm = (float*)e->v1_bm.p;
        vb.segsum = (float*)e->v1_segsum.p;
        vb.totblk = tblk;
        vb.cov = (float*)e->v1_cov.p;
        vb.msx = (float*)e->v1_msx.p;
        vb.xsum = (double*)e->v1_xsum.p;
        vb.cepcorr = (double*)e->v1_cepcorr.p;
        vb.cov3 = (double*)e->v1_cov3.p;
        vb.status = (int32_t*)e->v1_status.p;
        vb.ave = (double*)e->v1_ave.p;
        vb.sync5 = (double*)e->v1_sync5.p;
        vb.hasqi = hasqi ? 1 : 0;
        hb.cave = hasqi ? (double*)e->v1_cave.p : nullptr;
        e->last_launches += haspi_v1_run(g, hb, vb, cn, max_n24, e->f64, kt, s);
        e->last_launches += haspi_v1_finish(g, hb, vb, cn, (double*)e->out_haspi.p, (double*)e->out_raw.p, (int32_t*)e->out_hst.p, kt, s);
      }
    }
    // ---- ESTOI
    if (run_estoi) {
      RESERVE(e, e->x10, (size_t)2 * t10 * sizeof(float));
      RESERVE(e, e->st_energy, (size_t)std::max<int64_t>(tfr, 1) * sizeof(double));
      RESERVE(e, e->st_kept, (size_t)std::max<int64_t>(tfr, 1) * sizeof(int32_t));
      RESERVE(e, e->st_nkept, (size_t)cn * sizeof(int32_t));
      RESERVE(e, e->st_tob, (size_t)2 * std::max<int64_t>(tfr, 1) * 15 * sizeof(float));
      EstoiBuffers eb;
      memset(&eb, 0, sizeof(eb));
      eb.ref = d_ref;
      eb.deg = d_deg;
      eb.x10 = (float*)e->x10.p;
      eb.tot10 = t10;
      eb.energy = (double*)e->st_energy.p;
      eb.kept = (int32_t*)e->st_kept.p;
      eb.nkept = (int32_t*)e->st_nkept.p;
      eb.tob = (float*)e->st_tob.p;
      eb.totfr = tfr;
      eb.taps = (const double*)e->st_taps.p;
      eb.up = e->st_up;
      eb.down = e->st_down;
      eb.K = e->st_K;
      eb.score = (double*)e->out_estoi.p;
      eb.status = (int32_t*)e->out_est.p;
      e->last_launches += estoi_run(eg, eb, cn, max_n10, max_nfa, flags & NELE_FLAG_STOI_CLASSIC, kt, se);
    }
    // ---- SIIB main pipeline
    e->g_M.assign(cn, 0);
    e->g_offF.assign(cn, 0);
    e->g_F.assign(cn, 0);
    if (run_siib) {
      CU(e, cudaEventSynchronize(e->ev_M));
      int64_t tF = 0;
      const int64_t kMaxF = 400000;  // frames of one tiled signal (M is unbounded when almost nothing is active)
      for (int i = 0; i < cn; ++i) {
        const int M = e->h_M[i];
        e->g_M[i] = M;
        int64_t F = 0;
        if (M > 0) {
          const int64_t tl = std::max<int64_t>((int64_t)M * lens[first + i], 401);
          F = (tl - 400 + 199) / 200;
          if (F > kMaxF) F = 0;  // reported as too short
        }
        e->g_F[i] = F;
        e->g_offF[i] = tF;
        tF += F;
      }
      e->totF = std::max<int64_t>(tF, 1);
      GeomPacker sp;
      const size_t o_offF = sp.add(e->g_offF.data(), cn * sizeof(int64_t));
      const size_t o_F = sp.add(e->g_F.data(), cn * sizeof(int64_t));
      rc = push_blob(e, sp.blob, &e->h_sgeom, &e->h_sgeom_cap, e->sgeom, ss);
      if (rc != NELE_OK) return rc;
      sg.offF = (const int64_t*)((const char*)e->sgeom.p + o_offF);
      sg.F = (const int64_t*)((const char*)e->sgeom.p + o_F);
      const int64_t tFa = std::max<int64_t>(tF, 1);
      const bool siib_knn = flags & NELE_FLAG_SIIB_KNN;
      const int sub = std::min(cn, siib_knn ? 512 : kSiibSub);
      RESERVE(e, e->sb_mean, (size_t)cn * 2 * sizeof(double));
      RESERVE(e, e->sb_xdb, (size_t)tFa * sizeof(double));
      RESERVE(e, e->sb_act, (size_t)tFa * sizeof(int32_t));
      RESERVE(e, e->sb_aidx, (size_t)tFa * sizeof(int32_t));
      RESERVE(e, e->sb_src, (size_t)tFa * sizeof(int32_t));
      RESERVE(e, e->sb_Fa, (size_t)cn * sizeof(int32_t));
      RESERVE(e, e->sb_Pact, (size_t)cn * sizeof(int32_t));
      RESERVE(e, e->sb_perflag, (size_t)cn * 2 * sizeof(int32_t));
      RESERVE(e, e->sb_lograw, (size_t)2 * (tFa + 1) * 32 * sizeof(float));
      RESERVE(e, e->sb_logspec, (size_t)2 * (tFa + 1) * 32 * sizeof(float));
      RESERVE(e, e->sb_base, (size_t)sub * 59 * 1024 * sizeof(double));
      RESERVE(e, e->sb_Sxx, (size_t)sub * 420 * 420 * sizeof(double));
      RESERVE(e, e->sb_Sxy, (size_t)sub * 420 * 420 * sizeof(float));
      RESERVE(e, e->sb_Syy, (size_t)sub * 420 * 420 * sizeof(float));
      RESERVE(e, e->sb_Lc, (size_t)sub * 420 * 420 * sizeof(double));
      RESERVE(e, e->sb_G, (size_t)sub * 420 * 448 * sizeof(float));
      RESERVE(e, e->sb_perm, (size_t)sub * 420 * sizeof(int32_t));
      RESERVE(e, e->sb_rank, (size_t)cn * sizeof(int32_t));
      RESERVE(e, e->sb_sweeps, (size_t)cn * 17 * sizeof(int32_t));
      RESERVE(e, e->sb_lambda, (size_t)cn * 420 * sizeof(float));
      RESERVE(e, e->sb_rho, (size_t)cn * 420 * sizeof(float));
      RESERVE(e, e->sb_info, (size_t)cn * 8 * sizeof(double));
      sb.mean = (double*)e->sb_mean.p;
      sb.xdb = (double*)e->sb_xdb.p;
      sb.act = (int32_t*)e->sb_act.p;
      sb.aidx = (int32_t*)e->sb_aidx.p;
      sb.src = (int32_t*)e->sb_src.p;
      sb.Fa = (int32_t*)e->sb_Fa.p;
      sb.Pact = (int32_t*)e->sb_Pact.p;
      sb.perflag = (int32_t*)e->sb_perflag.p;
      static const bool no_proj = [] { const char* p = getenv("NELE_SIIB_QUADFORM"); return p && p[0] == '1'; }();
      sb.no_proj = no_proj ? 1 : 0;
      sb.lograw = (float*)e->sb_lograw.p;
      sb.logspec = (float*)e->sb_logspec.p;
      sb.totF = tFa + 1;
      sb.base = (double*)e->sb_base.p;
      sb.Sxx = (double*)e->sb_Sxx.p;
      sb.Sxy = (float*)e->sb_Sxy.p;
      sb.Syy = (float*)e->sb_Syy.p;
      sb.Lc = (double*)e->sb_Lc.p;
      sb.G = (float*)e->sb_G.p;
      sb.perm = (int32_t*)e->sb_perm.p;
      sb.rank = (int32_t*)e->sb_rank.p;
      sb.sweeps = (int32_t*)e->sb_sweeps.p;
      sb.sweep_rot = (int32_t*)e->sb_sweeps.p + cn;
      sb.lambda = (float*)e->sb_lambda.p;
      sb.rho = (float*)e->sb_rho.p;
      sb.info_part = (double*)e->sb_info.p;
      sb.score = (double*)e->out_siib.p;
      sb.status = (int32_t*)e->out_sst.p;
      CU(e, cudaMemsetAsync(e->sb_sweeps.p, 0, (size_t)cn * 17 * sizeof(int32_t), ss));
      static const bool use_jacobi = [] { const char* p = getenv("NELE_SIIB_JACOBI"); return p && p[0] == '1'; }();
      SiibEigBuffers egb;
      memset(&egb, 0, sizeof(egb));
      if (!use_jacobi) {
        RESERVE(e, e->eg_vec, (size_t)sub * 5 * 448 * sizeof(double));
        RESERVE(e, e->eg_zt, (size_t)sub * 420 * 448 * sizeof(float));
        double* ev = (double*)e->eg_vec.p;
        egb.d = ev;
        egb.e = ev + (size_t)sub * 448;
        egb.tau = ev + (size_t)sub * 448 * 2;
        egb.lam = ev + (size_t)sub * 448 * 3;
        egb.znorm = ev + (size_t)sub * 448 * 4;
        egb.zt = (float*)e->eg_zt.p;
        egb.scratch = (float*)e->sb_G.p;
        RESERVE(e, e->eg_gram, (size_t)sub * 112 * 112 * sizeof(double));
        egb.gram = (double*)e->eg_gram.p;
        RESERVE(e, e->eg_refl, (size_t)sub * 420 * 448 * sizeof(float));
        egb.refl = (float*)e->eg_refl.p;
      }
      SiibKnnBuffers kb;
      memset(&kb, 0, sizeof(kb));
      if (siib_knn) {
        int64_t maxF_all = 0;
        for (int i = 0; i < cn; ++i) maxF_all = std::max(maxF_all, e->g_F[i]);
        kb.ld = (std::max<int64_t>(maxF_all, 32) + 31) & ~(int64_t)31;
        constexpr int kNDig = 16384 + 2;
        if (!e->kn_digamma.p) {
          std::vector<double> dg(kNDig, 0.0);
          dg[1] = -0.57721566490153286;
          for (int m = 1; m + 1 < kNDig; ++m) dg[m + 1] = dg[m] + 1.0 / (double)m;
          RESERVE(e, e->kn_digamma, kNDig * sizeof(double));
          CU(e, cudaMemcpyAsync(e->kn_digamma.p, dg.data(), kNDig * sizeof(double), cudaMemcpyHostToDevice, ss));
          CU(e, cudaStreamSynchronize(ss));
        }
        RESERVE(e, e->kn_xk, (size_t)sub * 2 * 420 * kb.ld * sizeof(float));
        RESERVE(e, e->kn_info, (size_t)cn * 420 * sizeof(double));
        kb.xk = (float*)e->kn_xk.p;
        kb.info = (double*)e->kn_info.p;
        kb.digamma = (const double*)e->kn_digamma.p;
        kb.ndigamma = kNDig;
      }
      for (int lo_p = 0; lo_p < cn; lo_p += sub) {
        const int sn = std::min(sub, cn - lo_p);
        int64_t maxF = 0, maxU = 0;
        for (int i = 0; i < sn; ++i) {
          const int64_t Fi = e->g_F[lo_p + i], Li = lens[first + lo_p + i];
          maxF = std::max(maxF, Fi);
          maxU = std::max(maxU, std::min(Fi, Li / std::gcd<int64_t>(Li, 200)));  // distinct frames of the tiled signal
        }
        sb.pair_lo = lo_p;
        e->last_launches += siib_run(sg, sb, siib_knn ? &kb : nullptr, use_jacobi ? nullptr : &egb, sn, maxF, maxU, kt, ss);
        e->sub_lo = lo_p;
        e->sub_n = sn;
      }
    }
    if (!serial) {  // join
      CU(e, cudaEventRecord(e->ev_estoi, se));
      CU(e, cudaEventRecord(e->ev_siib, ss));
      CU(e, cudaStreamWaitEvent(s, e->ev_estoi, 0));
      CU(e, cudaStreamWaitEvent(s, e->ev_siib, 0));
    }
    CU(e, cudaEventRecord(e->ev1, s));
    CU(e, cudaGetLastError());
    if (trace) fprintf(stderr, "[nele] chunk %zu: all kernels queued at %.2f ms\n", ci, since());

    // ---- results
    std::vector<double> h_haspi(cn), h_raw((size_t)cn * kNumMod), h_estoi(cn), h_siib(cn);
    std::vector<int32_t> h_hst(cn), h_est(cn), h_sst(cn);
    if (run_haspi) {
      CU(e, cudaMemcpyAsync(h_haspi.data(), e->out_haspi.p, cn * sizeof(double), cudaMemcpyDeviceToHost, s));
      CU(e, cudaMemcpyAsync(h_raw.data(), e->out_raw.p, (size_t)cn * kNumMod * sizeof(double), cudaMemcpyDeviceToHost, s));
      CU(e, cudaMemcpyAsync(h_hst.data(), e->out_hst.p, cn * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    if (run_estoi) {
      CU(e, cudaMemcpyAsync(h_estoi.data(), e->out_estoi.p, cn * sizeof(double), cudaMemcpyDeviceToHost, s));
      CU(e, cudaMemcpyAsync(h_est.data(), e->out_est.p, cn * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    if (run_siib) {
      CU(e, cudaMemcpyAsync(h_siib.data(), e->out_siib.p, cn * sizeof(double), cudaMemcpyDeviceToHost, s));
      CU(e, cudaMemcpyAsync(h_sst.data(), e->out_sst.p, cn * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    CU(e, cudaStreamSynchronize(s));
    if (trace) fprintf(stderr, "[nele] chunk %zu: results on host at %.2f ms\n", ci, since());
    float ms = 0.f;
    CU(e, cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->last_kernel_ms += ms;
    if (kt) collect_kernel_times(e);
    for (int i = 0; i < cn; ++i) {
      const int gi = first + i;
      int32_t st = status ? status[gi] : 0;
      if (do_haspi) {
        int hs = NELE_ST_BAD_RATE;
        double v = kNaN;
        if (haspi_rate_ok) {
          hs = h_hst[i];
          v = h_haspi[i];
          if (mapped && hs == NELE_ST_OK && !haspi_v1) v = 1.0 / (1.0 + exp(-0.95 * (v - 2.8)));  // intel.py:116-120
          if (haspi_raw) memcpy(haspi_raw + (size_t)gi * kNumMod, h_raw.data() + (size_t)i * kNumMod, kNumMod * sizeof(double));
        } else if (haspi_raw) {
          for (int m = 0; m < kNumMod; ++m) haspi_raw[(size_t)gi * kNumMod + m] = kNaN;
        }
        scores[3 * gi + 1] = v;
        st = (st & ~0xff) | hs;
      }
      if (do_siib) {
        int ss = NELE_ST_BAD_RATE;
        double v = kNaN;
        if (siib_rate_ok) {
          ss = h_sst[i] & 0xff;
          st &= ~(int32_t)NELE_INFO_SIIB_NULLSPACE;
          if (h_sst[i] & 0x100) st |= (int32_t)NELE_INFO_SIIB_NULLSPACE;
          v = h_siib[i];
          if (mapped && ss == NELE_ST_OK) v = 1.0 / (1.0 + exp(-0.06 * (v - 32.0)));  // intel.py:102-106
        }
        scores[3 * gi + 0] = v;
        st = (st & ~0xff00) | (ss << 8);
      }
      if (do_estoi) {
        const int es = h_est[i];
        double v = h_estoi[i];
        if (mapped) v = 1.0 / (1.0 + exp(-8.0 * (v - 0.25)));  // intel.py:136-140 (the 1e-5 sentinel is mapped too)
        scores[3 * gi + 2] = v;
        st = (st & ~0xff0000) | (es << 16);
      }
      if (status) status[gi] = st;
    }
    e->stages_valid = (flags & NELE_FLAG_KEEP_STAGES) && plans.size() == 1;
    e->stage_v1 = haspi_v1;
    e->stage_metrics = (run_haspi ? NELE_METRIC_HASPI : 0) | (run_siib ? NELE_METRIC_SIIB : 0) | (run_estoi ? NELE_METRIC_ESTOI : 0);
  }
  return NELE_OK;
}

static int prefetch_core(nele_engine* e, const float* ref, const float* deg, const int64_t* offs, const int32_t* lens, int n,
                         uint32_t flags);

extern "C" int nele_prefetch(nele_engine* e, const float* ref, const float* deg, const int64_t* offs, const int32_t* lens,
                             int n, uint32_t flags) {
  if (!e) return NELE_E_ARG;
  std::lock_guard<std::mutex> lock(e->mu);
  e->pcm_noise = nullptr;
  return prefetch_core(e, ref, deg, offs, lens, n, flags);
}

extern "C" int nele_prefetch_pcm16(nele_engine* e, const int16_t* clean, const int16_t* enhanced, const int16_t* noise,
                                   const int64_t* offs, const int32_t* lens, int n, uint32_t flags) {
  if (!e) return NELE_E_ARG;
  std::lock_guard<std::mutex> lock(e->mu);
  if (!noise) return fail(e, NELE_E_ARG, "nele_prefetch_pcm16: null pointer");
  e->pcm_noise = noise;
  const int rc = prefetch_core(e, reinterpret_cast<const float*>(clean), reinterpret_cast<const float*>(enhanced), offs, lens, n, flags);
  e->pcm_noise = nullptr;
  return rc;
}

static int prefetch_core(nele_engine* e, const float* ref, const float* deg, const int64_t* offs, const int32_t* lens, int n,
                         uint32_t flags) {
  if (n <= 0 || !ref || !deg || !offs || !lens) return fail(e, NELE_E_ARG, "nele_prefetch: null pointer or n <= 0");
  if (flags & NELE_FLAG_DEVICE_INPUT) return NELE_OK;  // nothing to upload
  for (int i = 0; i < n; ++i)
    if (lens[i] <= 0 || offs[i] < 0) return fail(e, NELE_E_ARG, "nele_prefetch: pair %d has length %d / offset %lld", i, lens[i], (long long)offs[i]);
  CU(e, cudaSetDevice(e->device));
  const bool haspi_v1 = flags & NELE_FLAG_HASPI_V1;
  int max_pairs = haspi_v1 ? 2048 : 4096;
  if (const char* hc = getenv("NELE_HOST_CHUNK")) {
    const int v = atoi(hc);
    if (v > 0) max_pairs = std::min(v, max_pairs);
  }
  std::vector<ChunkPlan> plans;
  plan_chunks(offs, lens, n, false, max_pairs, haspi_v1 ? 64LL * 1000 * 1000 : 256LL * 1000 * 1000, plans);
  if (plans.size() != 1) return NELE_OK;  // multi-chunk calls pipeline their own uploads
  int slot = e->next_slot;
  if (e->pf[slot].valid) slot ^= 1;
  if (e->pf[slot].valid) return NELE_OK;  // both staging slots hold pending prefetches
  int rc = upload_chunk(e, plans[0], slot, ref, deg, offs, lens);
  if (rc != NELE_OK) return rc;
  e->pf[slot].valid = true;
  e->pf[slot].ref = ref;
  e->pf[slot].deg = deg;
  e->pf[slot].n = n;
  e->pf[slot].tot = plans[0].tot;
  e->pf[slot].seq = ++e->pf_seq;
  e->pf[slot].geom_hash = geom_hash(offs, lens, n);
  e->pf[slot].born = e->calls;
  e->pf[slot].pcm = e->pcm_noise != nullptr;
  return NELE_OK;
}

extern "C" int nele_prefetch_cancel(nele_engine* e) {
  if (!e) return NELE_E_ARG;
  std::lock_guard<std::mutex> lock(e->mu);
  e->pf[0].valid = e->pf[1].valid = false;
  return NELE_OK;
}

// ------------------------------------------------------------------ feature front-end
extern "C" int64_t nele_feature_frames(int32_t len) { return len < 0 ? 0 : 1 + (int64_t)len / 256; }

extern "C" int nele_features(nele_engine* e, const float* wav, const int64_t* offs, const int32_t* lens, int n,
                             uint32_t flags, double power, float* band, float* mag, float* phase, float* psd,
                             void* stream) {
  if (!e) return NELE_E_ARG;
  std::lock_guard<std::mutex> lock(e->mu);
  if (n < 0 || (n > 0 && (!wav || !offs || !lens || !band)))
    return fail(e, NELE_E_ARG, "nele_features: null pointer or negative n");
  if (flags & ~(NELE_FEAT_NOISE | NELE_FEAT_DEVICE_IO | NELE_FEAT_NO_POWER))
    return fail(e, NELE_E_ARG, "nele_features: bad flags 0x%x", flags);
  const bool noise = flags & NELE_FEAT_NOISE, dev_io = flags & NELE_FEAT_DEVICE_IO, normalize = !(flags & NELE_FEAT_NO_POWER);
  if (psd && !noise) return fail(e, NELE_E_ARG, "nele_features: psd requested without NELE_FEAT_NOISE");
  for (int i = 0; i < n; ++i)
    if (lens[i] <= 256 || offs[i] < 0)   // librosa: reflect padding by n_fft / 2 needs a longer signal
      return fail(e, NELE_E_ARG, "nele_features: waveform %d has length %d / offset %lld (need more than 256 samples)", i,
                  lens[i], (long long)offs[i]);
  e->last_kernel_ms = 0.0;
  e->last_launches = 0;
  e->kstats.clear();
  e->kt.count = 0;
  e->kt.enabled = e->profiling && e->kt_events;
  KernelTimer* kt = e->kt.enabled ? &e->kt : nullptr;
  if (n == 0) return NELE_OK;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;

  // frames per chunk: bounds the [257][T] workspaces at 4.3 GB each (NELE_FEAT_MAX_FRAMES: test knob for the chunk loop)
  static const int64_t kMaxFrames = [] {
    const char* p = getenv("NELE_FEAT_MAX_FRAMES");
    const long long v = p ? atoll(p) : 0;
    return (int64_t)(v > 0 ? v : (4 << 20));
  }();
  int64_t frame0 = 0;                   // frames of the waveforms before this chunk
  for (int first = 0; first < n;) {
    int last = first;
    int64_t frames = 0, lo = INT64_MAX, hi = 0;
    std::vector<int64_t> h_off, h_foff;
    std::vector<int2> h_tiles;
    while (last < n && (last == first || frames + nele_feature_frames(lens[last]) <= kMaxFrames)) {
      const int T = (int)nele_feature_frames(lens[last]);
      h_foff.push_back(frames);
      for (int t0 = 0; t0 < T; t0 += 8) h_tiles.push_back(make_int2(last - first, t0));
      frames += T;
      lo = std::min(lo, offs[last]);
      hi = std::max(hi, offs[last] + lens[last]);
      ++last;
    }
    const int cn = last - first;
    // geometry blob: offsets (relative to the staged span for host input), lengths, frame offsets, tiles
    const size_t g_off = 0, g_foff = g_off + 8 * (size_t)cn, g_tiles = g_foff + 8 * (size_t)cn,
                 g_len = g_tiles + 8 * h_tiles.size(), g_bytes = g_len + 4 * (size_t)cn;
    std::vector<char> blob(g_bytes);
    for (int i = 0; i < cn; ++i) h_off.push_back(dev_io ? offs[first + i] : offs[first + i] - lo);
    memcpy(&blob[g_off], h_off.data(), 8 * (size_t)cn);
    memcpy(&blob[g_foff], h_foff.data(), 8 * (size_t)cn);
    memcpy(&blob[g_tiles], h_tiles.data(), 8 * h_tiles.size());
    memcpy(&blob[g_len], lens + first, 4 * (size_t)cn);
    RESERVE(e, e->ft_geom, g_bytes);
    CU(e, cudaMemcpyAsync(e->ft_geom.p, blob.data(), g_bytes, cudaMemcpyHostToDevice, s));
    CU(e, cudaStreamSynchronize(s));   // blob is a local vector
    const char* gp = (const char*)e->ft_geom.p;

    const float* d_wav = wav;
    float *d_band = band + frame0 * 64, *d_mag = mag ? mag + frame0 * 257 : nullptr,
          *d_phase = phase ? phase + frame0 * 257 : nullptr, *d_psd = psd ? psd + frame0 * 257 : nullptr;
    const size_t mbytes = (size_t)frames * 257 * sizeof(float);
    if (!dev_io) {
      RESERVE(e, e->ft_wav, (size_t)(hi - lo) * sizeof(float));
      CU(e, cudaMemcpyAsync(e->ft_wav.p, wav + lo, (size_t)(hi - lo) * sizeof(float), cudaMemcpyHostToDevice, s));
      d_wav = (const float*)e->ft_wav.p;
      RESERVE(e, e->ft_band, (size_t)frames * 64 * sizeof(float));
      d_band = (float*)e->ft_band.p;
      if (mag || noise) {
        RESERVE(e, e->ft_mag, mbytes);
        d_mag = (float*)e->ft_mag.p;
      }
      if (phase) {
        RESERVE(e, e->ft_phase, mbytes);
        d_phase = (float*)e->ft_phase.p;
      }
      if (psd) {
        RESERVE(e, e->ft_psd, mbytes);
        d_psd = (float*)e->ft_psd.p;
      }
    } else if (noise && !d_mag) {
      RESERVE(e, e->ft_mag, mbytes);
      d_mag = (float*)e->ft_mag.p;
    }
    CU(e, cudaEventRecord(e->ev0, s));
    e->last_launches += features_run(d_wav, (const int64_t*)(gp + g_off), (const int32_t*)(gp + g_len),
                                     (const int64_t*)(gp + g_foff), (const int2*)(gp + g_tiles), cn, (int)h_tiles.size(),
                                     noise, (float)power, normalize, d_band, d_mag, d_phase, d_psd, kt, s);
    CU(e, cudaGetLastError());
    CU(e, cudaEventRecord(e->ev1, s));
    if (!dev_io) {
      CU(e, cudaMemcpyAsync(band + frame0 * 64, d_band, (size_t)frames * 64 * sizeof(float), cudaMemcpyDeviceToHost, s));
      if (mag) CU(e, cudaMemcpyAsync(mag + frame0 * 257, d_mag, mbytes, cudaMemcpyDeviceToHost, s));
      if (phase) CU(e, cudaMemcpyAsync(phase + frame0 * 257, d_phase, mbytes, cudaMemcpyDeviceToHost, s));
      if (psd) CU(e, cudaMemcpyAsync(psd + frame0 * 257, d_psd, mbytes, cudaMemcpyDeviceToHost, s));
    }
    CU(e, cudaStreamSynchronize(s));
    float ms = 0.f;
    CU(e, cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->last_kernel_ms += ms;
    collect_kernel_times(e);
    frame0 += frames;
    first = last;
  }
  return NELE_OK;
}

// ------------------------------------------------------------------ resynthesis (in-loop boundary)
extern "C" int nele_resyn(nele_engine* e, const float* clean, const float* noise, const int64_t* offs, const int32_t* lens,
                          int n, const float* alpha2, const int64_t* arow, uint32_t flags, float* enh, float* deg,
                          int32_t* out_lens, void* stream) {
  if (!e) return NELE_E_ARG;
  std::lock_guard<std::mutex> lock(e->mu);
  if (n < 0 || (n > 0 && (!clean || !offs || !lens || !alpha2 || (!enh && !deg) || (deg && !noise))))
    return fail(e, NELE_E_ARG, "nele_resyn: null pointer or negative n");
  if (flags & ~(NELE_RESYN_PCM16 | NELE_RESYN_ENH_ROUNDED)) return fail(e, NELE_E_ARG, "nele_resyn: bad flags 0x%x", flags);
  for (int i = 0; i < n; ++i)
    if (lens[i] <= 256 || offs[i] < 0)
      return fail(e, NELE_E_ARG, "nele_resyn: waveform %d has length %d / offset %lld (need more than 256 samples)", i, lens[i],
                  (long long)offs[i]);
  e->last_kernel_ms = 0.0;
  e->last_launches = 0;
  e->kstats.clear();
  e->kt.count = 0;
  e->kt.enabled = e->profiling && e->kt_events;
  KernelTimer* kt = e->kt.enabled ? &e->kt : nullptr;
  if (n == 0) return NELE_OK;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  std::vector<int64_t> h_foff(n);
  std::vector<int2> h_tiles;
  int64_t frames = 0;
  for (int i = 0; i < n; ++i) {
    const int T = (int)nele_feature_frames(lens[i]);
    h_foff[i] = arow ? arow[i] : frames;
    for (int t0 = 0; t0 < T - 1; t0 += 7) h_tiles.push_back(make_int2(i, t0));   // 8 frames -> 7 output hops per CTA
    frames += T;
    if (out_lens) out_lens[i] = 256 * (lens[i] / 256);   // len(librosa.istft(...)) = hop * (T - 1); audio_util.py:190-193 trims to it
  }
  const size_t g_off = 0, g_foff = g_off + 8 * (size_t)n, g_tiles = g_foff + 8 * (size_t)n,
               g_len = g_tiles + 8 * h_tiles.size(), g_bytes = g_len + 4 * (size_t)n;
  std::vector<char> blob(g_bytes);
  memcpy(&blob[g_off], offs, 8 * (size_t)n);
  memcpy(&blob[g_foff], h_foff.data(), 8 * (size_t)n);
  memcpy(&blob[g_tiles], h_tiles.data(), 8 * h_tiles.size());
  memcpy(&blob[g_len], lens, 4 * (size_t)n);
  RESERVE(e, e->ft_geom, g_bytes);
  CU(e, cudaMemcpyAsync(e->ft_geom.p, blob.data(), g_bytes, cudaMemcpyHostToDevice, s));
  CU(e, cudaStreamSynchronize(s));   // blob is a local vector
  const char* gp = (const char*)e->ft_geom.p;
  CU(e, cudaEventRecord(e->ev0, s));
  e->last_launches += resyn_run(clean, noise, (const int64_t*)(gp + g_off), (const int32_t*)(gp + g_len),
                                (const int64_t*)(gp + g_foff), (const int2*)(gp + g_tiles), (int)h_tiles.size(), alpha2,
                                (int)(flags & (NELE_RESYN_PCM16 | NELE_RESYN_ENH_ROUNDED)), enh, deg, kt, s);
  CU(e, cudaGetLastError());
  CU(e, cudaEventRecord(e->ev1, s));
  CU(e, cudaStreamSynchronize(s));
  float ms = 0.f;
  CU(e, cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->last_kernel_ms += ms;
  collect_kernel_times(e);
  return NELE_OK;
}

extern "C" int nele_last_timing(const nele_engine* e, double* kernel_ms, int64_t* launches) {
  if (!e) return NELE_E_ARG;
  if (kernel_ms) *kernel_ms = e->last_kernel_ms;
  if (launches) *launches = e->last_launches;
  return NELE_OK;
}

extern "C" int nele_kernel_time(const nele_engine* e, int idx, const char** name, double* ms, int64_t* launches) {
  if (!e || idx < 0 || idx >= (int)e->kstats.size()) return NELE_E_ARG;
  if (name) *name = e->kstats[idx].name.c_str();
  if (ms) *ms = e->kstats[idx].ms;
  if (launches) *launches = e->kstats[idx].launches;
  return NELE_OK;
}

extern "C" int nele_get_stage(nele_engine* e, const char* name, int pair, void* dst, size_t cap, size_t* nbytes) {
  if (!e || !name) return NELE_E_ARG;
  std::lock_guard<std::mutex> lock(e->mu);
  if (!e->stages_valid) return fail(e, NELE_E_ARG, "nele_get_stage: no stages kept (call nele_score_batch with NELE_FLAG_KEEP_STAGES on a single-chunk batch)");
  if (pair < 0 || pair >= e->chunk_n) return fail(e, NELE_E_ARG, "nele_get_stage: pair %d out of range", pair);
  CU(e, cudaSetDevice(e->device));
  struct Piece { const void* src; size_t bytes; };
  std::vector<Piece> pieces;
  std::vector<int32_t> ints;  // host-side values returned as-is
  const int64_t o24 = e->g_off24[pair], osub = e->g_offsub[pair];
  const int n24 = e->g_n24[pair], nsub = e->g_nsub[pair];
  std::string nm(name);
  const uint32_t need = (nm.rfind("haspi.", 0) == 0 || nm.rfind("haspi1.", 0) == 0) ? NELE_METRIC_HASPI : nm.rfind("estoi.", 0) == 0 ? NELE_METRIC_ESTOI : NELE_METRIC_SIIB;
  if (!(e->stage_metrics & need)) return fail(e, NELE_E_ARG, "nele_get_stage: '%s' belongs to a metric the last call did not run", name);
  auto dev_i32 = [&](const DevBuf& b, size_t idx, int32_t* out) -> cudaError_t {
    return cudaMemcpy(out, (const int32_t*)b.p + idx, sizeof(int32_t), cudaMemcpyDeviceToHost);
  };
  if (e->stage_v1 && (nm == "haspi.envlp" || nm == "haspi.nsel" || nm == "haspi.cep" || nm == "haspi.cepmean"))
    return fail(e, NELE_E_ARG, "nele_get_stage: '%s' is a HASPI version 2 stage; the last call ran version 1", name);
  if (nm == "haspi.mid") {
    for (int q = 0; q < 2; ++q) pieces.push_back({(float*)e->mid.p + q * e->tot24 + o24, n24 * sizeof(float)});
  } else if (nm == "haspi.x24") {
    for (int q = 0; q < 2; ++q) pieces.push_back({(float*)e->x24.p + q * e->tot24 + o24, n24 * sizeof(float)});
  } else if (nm == "haspi.bw") {
    pieces.push_back({(double*)e->bw.p + (size_t)pair * 2 * kBands, 2 * kBands * sizeof(double)});
  } else if (nm == "haspi.shift") {
    pieces.push_back({(int32_t*)e->shift.p + (size_t)pair * kBands, kBands * sizeof(int32_t)});
  } else if (nm == "haspi.envlp") {
    for (int q = 0; q < 2; ++q) pieces.push_back({(float*)e->envlp.p + (q * e->totsub + osub) * kBands, (size_t)nsub * kBands * sizeof(float)});
  } else if (nm == "haspi.nsel") {
    pieces.push_back({(int32_t*)e->nsel.p + pair, sizeof(int32_t)});
  } else if (nm == "haspi.cep") {
    int32_t nsel = 0;
    CU(e, dev_i32(e->nsel, pair, &nsel));
    for (int q = 0; q < 2; ++q)
      for (int j = 0; j < kNumCep; ++j)
        pieces.push_back({(float*)e->cep.p + (size_t)(q * kNumCep + j) * e->totsub + osub, (size_t)nsel * sizeof(float)});
  } else if (nm == "haspi.cepmean") {
    pieces.push_back({(double*)e->cepmean.p + (size_t)pair * 2 * kNumCep, 2 * kNumCep * sizeof(double)});
  } else if (nm == "haspi1.segsum" || nm == "haspi1.cov" || nm == "haspi1.msx") {
    if (!e->stage_v1) return fail(e, NELE_E_ARG, "nele_get_stage: '%s' needs a NELE_FLAG_HASPI_V1 call", name);
    const int64_t ob = e->g_offblk[pair];
    const int nblk = (n24 + 191) / 192;
    const int nseg = n24 < 192 ? 0 : 1 + n24 / 384 + (n24 - 192) / 384;
    if (nm == "haspi1.segsum") {
      for (int q = 0; q < 4; ++q) pieces.push_back({(float*)e->v1_segsum.p + (q * e->totblk + ob) * kBands, (size_t)nblk * kBands * sizeof(float)});
    } else {
      pieces.push_back({(float*)(nm == "haspi1.cov" ? e->v1_cov.p : e->v1_msx.p) + ob * kBands, (size_t)nseg * kBands * sizeof(float)});
    }
  } else if (nm == "estoi.x10") {
    for (int q = 0; q < 2; ++q) pieces.push_back({(float*)e->x10.p + q * e->tot10 + e->g_off10[pair], (size_t)e->g_n10[pair] * sizeof(float)});
  } else if (nm == "estoi.info") {
    int32_t nk = 0;
    CU(e, dev_i32(e->st_nkept, pair, &nk));
    ints = {e->g_n10[pair], e->g_nfa[pair], nk};
  } else if (nm == "estoi.kept") {
    int32_t nk = 0;
    CU(e, dev_i32(e->st_nkept, pair, &nk));
    pieces.push_back({(int32_t*)e->st_kept.p + e->g_offfr[pair], (size_t)nk * sizeof(int32_t)});
  } else if (nm == "estoi.tob") {
    int32_t nk = 0;
    CU(e, dev_i32(e->st_nkept, pair, &nk));
    const int nfr = std::max(nk - 1, 0);
    for (int q = 0; q < 2; ++q) pieces.push_back({(float*)e->st_tob.p + (q * e->totfr + e->g_offfr[pair]) * 15, (size_t)nfr * 15 * sizeof(float)});
  } else if (nm == "siib.tile") {
    int32_t wa = 0, fa = 0;
    CU(e, dev_i32(e->sb_wact, pair, &wa));
    CU(e, dev_i32(e->sb_Fa, pair, &fa));
    ints = {e->g_M[pair], wa, (int32_t)e->g_F[pair], fa};
  } else if (nm == "siib.logspec") {
    int32_t fa = 0;
    CU(e, dev_i32(e->sb_Fa, pair, &fa));
    for (int q = 0; q < 2; ++q) pieces.push_back({(float*)e->sb_logspec.p + (q * (e->totF + 1) + e->g_offF[pair]) * 32, (size_t)fa * 32 * sizeof(float)});
  } else if (nm == "siib.lambda") {
    pieces.push_back({(float*)e->sb_lambda.p + (size_t)pair * 420, 420 * sizeof(float)});
  } else if (nm == "siib.rho") {
    pieces.push_back({(float*)e->sb_rho.p + (size_t)pair * 420, 420 * sizeof(float)});
  } else if (nm == "siib.rank") {
    int32_t rk = 0, sw = 0;
    CU(e, dev_i32(e->sb_rank, pair, &rk));
    CU(e, dev_i32(e->sb_sweeps, pair, &sw));
    ints = {rk, sw};
    for (int k = 0; k < 16; ++k) {
      int32_t v = 0;
      CU(e, dev_i32(e->sb_sweeps, (size_t)e->chunk_n + (size_t)pair * 16 + k, &v));
      ints.push_back(v);
    }
  } else if (nm == "siib.sxx" || nm == "siib.sxy" || nm == "siib.syy") {
    if (pair < e->sub_lo || pair >= e->sub_lo + e->sub_n) return fail(e, NELE_E_ARG, "nele_get_stage: '%s' is only kept for the last SIIB sub-chunk", name);
    const size_t lp = pair - e->sub_lo;
    if (nm == "siib.sxx") pieces.push_back({(double*)e->sb_Sxx.p + lp * 420 * 420, (size_t)420 * 420 * sizeof(double)});
    else if (nm == "siib.sxy") pieces.push_back({(float*)e->sb_Sxy.p + lp * 420 * 420, (size_t)420 * 420 * sizeof(float)});
    else pieces.push_back({(float*)e->sb_Syy.p + lp * 420 * 420, (size_t)420 * 420 * sizeof(float)});
  } else {
    return fail(e, NELE_E_ARG, "nele_get_stage: unknown stage '%s'", name);
  }
  size_t total = ints.size() * sizeof(int32_t);
  for (auto& p : pieces) total += p.bytes;
  if (nbytes) *nbytes = total;
  if (!dst) return NELE_OK;
  if (cap < total) return fail(e, NELE_E_ARG, "nele_get_stage: buffer too small (%zu < %zu)", cap, total);
  char* d = (char*)dst;
  if (!ints.empty()) {
    memcpy(d, ints.data(), ints.size() * sizeof(int32_t));
    d += ints.size() * sizeof(int32_t);
  }
  for (auto& p : pieces) {
    if (p.bytes) CU(e, cudaMemcpy(d, p.src, p.bytes, cudaMemcpyDeviceToHost));
    d += p.bytes;
  }
  return NELE_OK;
}
