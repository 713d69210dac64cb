// libnele_score.so -- C ABI (include/nele_score.h) and host orchestration.
//
// One engine per (process, device).  A call splits the batch into sub-batches
// ("chunks") that bound the device workspace, stages the chunk's waveforms in
// HBM, runs the metric pipelines on one stream and copies the per-pair records
// back.  There is no CPU implementation of any metric in this library: without
// a CUDA device nele_create fails.
#include "../../include/nele_score.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
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

struct nele_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  bool f64 = false;  // recurrence precision of the ear model (NELE_HASPI_F64=1)

  // constant tables
  DevBuf bands, rs_taps;
  int rs_fs = 0, rs_up = 1, rs_down = 1;
  double hl_cached[6] = {-1, -1, -1, -1, -1, -1};

  // workspace (grow-only)
  DevBuf in_ref, in_deg, geom, x24, mid, bw, shift, envlp, rowsel, nsel, cep, cepmean, modsum, dither;
  DevBuf out_intel, out_raw, out_status;
  DevBuf estoi_ws, siib_ws;

  // geometry of the last chunk (for nele_get_stage)
  bool stages_valid = false;
  std::vector<int64_t> g_off16, g_off24, g_offsub;
  std::vector<int32_t> g_len16, g_n24, g_nsub;
  int64_t tot24 = 0, totsub = 0;
  int chunk_n = 0;
  std::vector<int32_t> h_nsel;

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
  haspi_upload_constants(e->stream);
  {
    float cepm[kBands * kNumCep];
    host::make_cep_basis(cepm);
    host::ModFilters mf;
    host::make_mod_filters(mf);
    if ((int)mf.taps.size() != 2850) {
      fail(nullptr, NELE_E_ARG, "modulation filter design produced %zu taps, expected 2850", mf.taps.size());
      delete e;
      return NELE_E_ARG;
    }
    haspi_upload_tables(cepm, mf.nhalf, mf.offset, mf.taps.data(), (int)mf.taps.size(), e->stream);
  }
  CUC(cudaGetLastError());
#undef CUC
  *out = e;
  return NELE_OK;
}

extern "C" void nele_destroy(nele_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  DevBuf* all[] = {&e->bands, &e->rs_taps, &e->in_ref, &e->in_deg, &e->geom, &e->x24, &e->mid, &e->bw, &e->shift,
                   &e->envlp, &e->rowsel, &e->nsel, &e->cep, &e->cepmean, &e->modsum, &e->dither, &e->out_intel,
                   &e->out_raw, &e->out_status, &e->estoi_ws, &e->siib_ws};
  for (DevBuf* b : all)
    if (b->p) cudaFree(b->p);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

static int ensure_tables(nele_engine* e, int fs, const double* hl, cudaStream_t s) {
  double h[6] = {0, 0, 0, 0, 0, 0};
  if (hl) memcpy(h, hl, sizeof(h));
  if (!e->bands.p || memcmp(h, e->hl_cached, sizeof(h)) != 0) {
    BandConst bc[kBands];
    host::make_band_consts(h, bc);
    RESERVE(e, e->bands, sizeof(bc));
    CU(e, cudaMemcpyAsync(e->bands.p, bc, sizeof(bc), cudaMemcpyHostToDevice, s));
    CU(e, cudaStreamSynchronize(s));
    memcpy(e->hl_cached, h, sizeof(h));
  }
  if (e->rs_fs != fs) {
    host::ResampyTaps rt;
    if (fs == kFs24) {
      rt.up = rt.down = 1;
      rt.taps.assign(128, 0.0);
    } else {
      host::make_resampy_taps(fs, kFs24, rt);
    }
    RESERVE(e, e->rs_taps, rt.taps.size() * sizeof(double));
    CU(e, cudaMemcpyAsync(e->rs_taps.p, rt.taps.data(), rt.taps.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(e, cudaStreamSynchronize(s));
    e->rs_fs = fs;
    e->rs_up = rt.up;
    e->rs_down = rt.down;
  }
  return NELE_OK;
}

static const double kNaN = nan("");

extern "C" int nele_score_batch(nele_engine* e, const float* ref, const float* deg, const int64_t* offs,
                                const int32_t* lens, int n, int fs, uint32_t metrics, uint32_t flags,
                                const float* dither, int64_t dither_rows, uint64_t seed, const double* hl,
                                double* scores, double* haspi_raw, int32_t* status, void* stream) {
  if (!e) return NELE_E_ARG;
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
  if (n == 0) return NELE_OK;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
  const bool do_haspi = metrics & NELE_METRIC_HASPI, do_siib = metrics & NELE_METRIC_SIIB, do_estoi = metrics & NELE_METRIC_ESTOI;
  const bool haspi_rate_ok = fs <= kFs24;
  const bool dev_in = flags & NELE_FLAG_DEVICE_INPUT;
  const bool mapped = flags & NELE_FLAG_MAPPED;
  int rc = ensure_tables(e, haspi_rate_ok ? fs : kFs24, hl, s);
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

  // ---- chunking: bound the workspace by pairs and by total samples
  const int64_t kMaxChunkSamples = 80LL * 1000 * 1000;  // input-rate samples per signal
  const int kMaxChunkPairs = 2048;
  int first = 0;
  while (first < n) {
    int last = first;
    int64_t tot = 0, lo = offs[first], hi = offs[first] + lens[first];
    while (last < n && last - first < kMaxChunkPairs && (last == first || tot + lens[last] <= kMaxChunkSamples)) {
      tot += lens[last];
      lo = std::min(lo, offs[last]);
      hi = std::max(hi, offs[last] + (int64_t)lens[last]);
      ++last;
    }
    const int cn = last - first;
    // geometry
    e->g_off16.resize(cn); e->g_len16.resize(cn); e->g_off24.resize(cn); e->g_n24.resize(cn);
    e->g_offsub.resize(cn); e->g_nsub.resize(cn);
    int64_t t24 = 0, tsub = 0;
    int max_nsub = 0;
    const bool span_copy = !dev_in && (hi - lo) <= 2 * tot + 4096;
    int64_t packed = 0;
    for (int i = 0; i < cn; ++i) {
      const int L = lens[first + i];
      e->g_len16[i] = L;
      e->g_off16[i] = dev_in ? offs[first + i] : (span_copy ? offs[first + i] - lo : packed);
      packed += (L + 3) & ~3;
      const int n24 = (fs == kFs24 || !haspi_rate_ok) ? L : (int)(((int64_t)L * e->rs_up + e->rs_down - 1) / e->rs_down);
      const int nsub = (n24 + kDecim - 1) / kDecim;
      e->g_n24[i] = n24;
      e->g_nsub[i] = nsub;
      e->g_off24[i] = t24;
      e->g_offsub[i] = tsub;
      t24 += (n24 + 31) & ~31;
      tsub += nsub;
      max_nsub = std::max(max_nsub, nsub);
    }
    e->tot24 = t24;
    e->totsub = tsub;
    e->chunk_n = cn;

    // inputs
    const float *d_ref = ref, *d_deg = deg;
    if (!dev_in) {
      const size_t in_elems = span_copy ? (size_t)(hi - lo) : (size_t)packed;
      RESERVE(e, e->in_ref, in_elems * sizeof(float));
      RESERVE(e, e->in_deg, in_elems * sizeof(float));
      if (span_copy) {
        CU(e, cudaMemcpyAsync(e->in_ref.p, ref + lo, in_elems * sizeof(float), cudaMemcpyHostToDevice, s));
        CU(e, cudaMemcpyAsync(e->in_deg.p, deg + lo, in_elems * sizeof(float), cudaMemcpyHostToDevice, s));
      } else {
        for (int i = 0; i < cn; ++i) {
          CU(e, cudaMemcpyAsync((float*)e->in_ref.p + e->g_off16[i], ref + offs[first + i], sizeof(float) * lens[first + i], cudaMemcpyHostToDevice, s));
          CU(e, cudaMemcpyAsync((float*)e->in_deg.p + e->g_off16[i], deg + offs[first + i], sizeof(float) * lens[first + i], cudaMemcpyHostToDevice, s));
        }
      }
      d_ref = (const float*)e->in_ref.p;
      d_deg = (const float*)e->in_deg.p;
    }
    // geometry arrays -> device (one buffer)
    const size_t gbytes = (size_t)cn * (3 * sizeof(int64_t) + 3 * sizeof(int32_t));
    RESERVE(e, e->geom, gbytes + 64);
    char* gp = (char*)e->geom.p;
    PairGeom g;
    g.off16 = (const int64_t*)gp;
    g.off24 = g.off16 + cn;
    g.offsub = g.off24 + cn;
    g.len16 = (const int32_t*)(g.offsub + cn);
    g.n24 = g.len16 + cn;
    g.nsub = g.n24 + cn;
    CU(e, cudaMemcpyAsync((void*)g.off16, e->g_off16.data(), cn * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    CU(e, cudaMemcpyAsync((void*)g.off24, e->g_off24.data(), cn * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    CU(e, cudaMemcpyAsync((void*)g.offsub, e->g_offsub.data(), cn * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    CU(e, cudaMemcpyAsync((void*)g.len16, e->g_len16.data(), cn * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU(e, cudaMemcpyAsync((void*)g.n24, e->g_n24.data(), cn * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU(e, cudaMemcpyAsync((void*)g.nsub, e->g_nsub.data(), cn * sizeof(int32_t), cudaMemcpyHostToDevice, s));

    RESERVE(e, e->out_intel, cn * sizeof(double));
    RESERVE(e, e->out_raw, (size_t)cn * kNumMod * sizeof(double));
    RESERVE(e, e->out_status, (size_t)cn * 3 * sizeof(int32_t));

    CU(e, cudaEventRecord(e->ev0, s));
    HaspiBuffers hb;
    memset(&hb, 0, sizeof(hb));
    if (do_haspi && haspi_rate_ok) {
      RESERVE(e, e->x24, (size_t)2 * t24 * sizeof(float));
      RESERVE(e, e->mid, (size_t)2 * t24 * sizeof(double));
      RESERVE(e, e->bw, (size_t)cn * 2 * kBands * sizeof(double));
      RESERVE(e, e->shift, (size_t)cn * kBands * sizeof(int32_t));
      RESERVE(e, e->envlp, (size_t)2 * tsub * kBands * sizeof(float));
      RESERVE(e, e->rowsel, (size_t)tsub * sizeof(int32_t));
      RESERVE(e, e->nsel, (size_t)cn * sizeof(int32_t));
      RESERVE(e, e->cep, (size_t)2 * kNumCep * tsub * sizeof(float));
      RESERVE(e, e->cepmean, (size_t)cn * 2 * kNumCep * sizeof(double));
      RESERVE(e, e->modsum, (size_t)cn * kNumCep * kNumMod * 5 * sizeof(double));
      hb.ref = d_ref;
      hb.deg = d_deg;
      hb.x24 = (float*)e->x24.p;
      hb.mid = (double*)e->mid.p;
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
      e->last_launches += haspi_run(g, hb, cn, max_nsub, e->f64, s);
      e->last_launches += haspi_finish(hb, cn, (double*)e->out_intel.p, (double*)e->out_raw.p, (int32_t*)e->out_status.p, s);
    }
    CU(e, cudaEventRecord(e->ev1, s));
    CU(e, cudaGetLastError());

    // results
    std::vector<double> h_intel(cn), h_raw((size_t)cn * kNumMod);
    std::vector<int32_t> h_st((size_t)cn * 3);
    if (do_haspi && haspi_rate_ok) {
      CU(e, cudaMemcpyAsync(h_intel.data(), e->out_intel.p, cn * sizeof(double), cudaMemcpyDeviceToHost, s));
      CU(e, cudaMemcpyAsync(h_raw.data(), e->out_raw.p, (size_t)cn * kNumMod * sizeof(double), cudaMemcpyDeviceToHost, s));
      CU(e, cudaMemcpyAsync(h_st.data(), e->out_status.p, cn * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    CU(e, cudaStreamSynchronize(s));
    float ms = 0.f;
    CU(e, cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->last_kernel_ms += ms;
    for (int i = 0; i < cn; ++i) {
      const int gi = first + i;
      int32_t st = status ? status[gi] : 0;
      if (do_haspi) {
        int hs;
        double v;
        if (!haspi_rate_ok) {
          hs = NELE_ST_BAD_RATE;
          v = kNaN;
        } else {
          hs = h_st[i];
          v = h_intel[i];
          if (mapped && hs == NELE_ST_OK) v = 1.0 / (1.0 + exp(-0.95 * (v - 2.8)));  // intel.py:116-120
          if (haspi_raw) memcpy(haspi_raw + (size_t)gi * kNumMod, h_raw.data() + (size_t)i * kNumMod, kNumMod * sizeof(double));
        }
        scores[3 * gi + 1] = v;
        st = (st & ~0xff) | hs;
      }
      (void)do_siib;
      (void)do_estoi;
      if (status) status[gi] = st;
    }
    e->stages_valid = (flags & NELE_FLAG_KEEP_STAGES) && first == 0 && last == n;
    first = last;
  }
  return NELE_OK;
}

extern "C" int nele_last_timing(const nele_engine* e, double* kernel_ms, int64_t* launches) {
  if (!e) return NELE_E_ARG;
  if (kernel_ms) *kernel_ms = e->last_kernel_ms;
  if (launches) *launches = e->last_launches;
  return NELE_OK;
}

extern "C" int nele_get_stage(nele_engine* e, const char* name, int pair, void* dst, size_t cap, size_t* nbytes) {
  if (!e || !name) return NELE_E_ARG;
  if (!e->stages_valid) return fail(e, NELE_E_ARG, "nele_get_stage: no stages kept (call nele_score_batch with NELE_FLAG_KEEP_STAGES on a single-chunk batch)");
  if (pair < 0 || pair >= e->chunk_n) return fail(e, NELE_E_ARG, "nele_get_stage: pair %d out of range", pair);
  CU(e, cudaSetDevice(e->device));
  struct Piece { const void* src; size_t bytes; };
  std::vector<Piece> pieces;
  const int64_t o24 = e->g_off24[pair], osub = e->g_offsub[pair];
  const int n24 = e->g_n24[pair], nsub = e->g_nsub[pair];
  std::string nm(name);
  int32_t nsel = 0;
  if (nm == "haspi.cep" || nm == "haspi.nsel") {
    CU(e, cudaMemcpy(&nsel, (int32_t*)e->nsel.p + pair, sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  if (nm == "haspi.mid") {
    for (int q = 0; q < 2; ++q) pieces.push_back({(double*)e->mid.p + q * e->tot24 + o24, n24 * sizeof(double)});
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
    for (int q = 0; q < 2; ++q)
      for (int j = 0; j < kNumCep; ++j)
        pieces.push_back({(float*)e->cep.p + (size_t)(q * kNumCep + j) * e->totsub + osub, (size_t)nsel * sizeof(float)});
  } else if (nm == "haspi.cepmean") {
    pieces.push_back({(double*)e->cepmean.p + (size_t)pair * 2 * kNumCep, 2 * kNumCep * sizeof(double)});
  } else {
    return fail(e, NELE_E_ARG, "nele_get_stage: unknown stage '%s'", name);
  }
  size_t total = 0;
  for (auto& p : pieces) total += p.bytes;
  if (nbytes) *nbytes = total;
  if (!dst) return NELE_OK;
  if (cap < total) return fail(e, NELE_E_ARG, "nele_get_stage: buffer too small (%zu < %zu)", cap, total);
  char* d = (char*)dst;
  for (auto& p : pieces) {
    CU(e, cudaMemcpy(d, p.src, p.bytes, cudaMemcpyDeviceToHost));
    d += p.bytes;
  }
  return NELE_OK;
}
