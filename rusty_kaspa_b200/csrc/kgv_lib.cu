// kgv_lib.cu — kernels and C ABI of libkgv.so (see include/kgv.h).
//
// Hand-written CUDA for sm_100a.  No CPU fallback: every entry point needs a CUDA device.
#include "../../include/kgv.h"
#include "kgv_internal.h"
#include "kgv_verify.cuh"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

using namespace kgv;

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
// per-thread table in shared memory, word-major / thread-minor: every access of a warp hits 32
// consecutive banks whatever entry each lane selects
struct SmemTab {
  uint32_t* base;  // smem + threadIdx.x
  __device__ __forceinline__ void put(int e, int w, uint32_t v) { base[(e * 16 + w) * KGV_BLOCK] = v; }
  __device__ __forceinline__ uint32_t get(int e, int w) const { return base[(e * 16 + w) * KGV_BLOCK]; }
};

// 256-bit read-only load (LDG.E.256 on sm_100a)
__device__ __forceinline__ void ldg256(uint32_t* w, const void* p) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
// streaming variant for the signature triples (read once)
__device__ __forceinline__ void ldg256_stream(uint32_t* w, const void* p) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}

// generator table entry: 64 bytes, 64-byte aligned: two 256-bit loads
struct GLoadDev {
  __device__ __forceinline__ void operator()(fe& x, fe& y, const uint32_t* entry) const {
    ldg256(x.v, entry);
    ldg256(y.v, entry + 8);
  }
};

// 32 big-endian bytes -> 8 numeric words (w[0] most significant)
template <bool ALIGNED>
__device__ __forceinline__ void load_be32(uint32_t* w, const uint8_t* p) {
  if (ALIGNED) {
    ldg256_stream(w, p);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = bswap32(w[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++)
      w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | (uint32_t)p[4 * i + 3];
  }
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_build_gtab(uint32_t* __restrict__ gtab) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2u * 65536u) return;
  uint32_t v = t & 0xFFFFu;
  uint32_t which = t >> 16;
  uint32_t* out = gtab + (size_t)t * 16;
  if (v == 0) {
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = 0;
    return;
  }
  const fe gx = {KGV_GX_LIMBS}, gy = {KGV_GY_LIMBS}, hx = {KGV_G128X_LIMBS}, hy = {KGV_G128Y_LIMBS};
  fe x, y;
  if (which == 0) gtab_entry(x, y, v, gx, gy);
  else gtab_entry(x, y, v, hx, hy);
#pragma unroll
  for (int i = 0; i < 8; i++) { out[i] = x.v[i]; out[8 + i] = y.v[i]; }
}

// Each thread verifies KGV_ITEMS consecutive-stride items (i = tid + j * total_threads: coalesced) and shares
// ONE modular inversion among them (Montgomery's trick): the field inversion of BIP-340's final affine
// conversion, resp. the scalar inversion s^-1 of ECDSA, drops from 1 to 1/KGV_ITEMS per signature with no
// cross-thread synchronisation.  Pending state sits in (L1-resident) local memory between the phases.
template <bool ALIGNED, bool INDEXED>
__global__ void __launch_bounds__(KGV_BLOCK, KGV_BLOCKS_PER_SM)
k_schnorr_verify(const uint8_t* __restrict__ pk, const uint8_t* __restrict__ msg, const uint8_t* __restrict__ sig, size_t n_arg,
                 uint8_t* __restrict__ status, const uint32_t* __restrict__ gtab, const uint32_t* __restrict__ index, const uint32_t* __restrict__ n_dev) {
  // INDEXED: verify only the listed items (the signature-cache misses), their count read on the device.  A separate instantiation: the two
  // extra pointers live across the whole kernel cost the plain form 3 % (register pressure at the 168-register cap, measured).
  extern __shared__ uint32_t smem[];
  const size_t n = (INDEXED && n_dev) ? (size_t)*n_dev : n_arg;
  const size_t total = (size_t)gridDim.x * KGV_BLOCK;
  const size_t tid = (size_t)blockIdx.x * KGV_BLOCK + threadIdx.x;
  SmemTab tab{smem + threadIdx.x};
  fe X[KGV_ITEMS], Y[KGV_ITEMS], ZT[KGV_ITEMS], RX[KGV_ITEMS], pre[KGV_ITEMS];
  uint8_t st[KGV_ITEMS];
  // persistent grid (one resident wave): every thread walks the batch with stride total*KGV_ITEMS, so all
  // SM slots finish within one item of each other whatever n is (no wave quantisation)
#pragma unroll 1
  for (size_t base = 0; base < n; base += total * KGV_ITEMS) {
  fe acc;
  fe_set_u32(acc, 1);
  bool any = false;
#pragma unroll 1
  for (int j = 0; j < KGV_ITEMS; j++) {
    size_t i = base + tid + (size_t)j * total;
    st[j] = KGV_ST_INVALID;
    if (i >= n) continue;
    if (INDEXED) i = index[i];
    uint32_t pkw[8], mw[8], sw[16];
    load_be32<ALIGNED>(pkw, pk + 32 * i);
    load_be32<ALIGNED>(mw, msg + 32 * i);
    load_be32<ALIGNED>(sw, sig + 64 * i);
    load_be32<ALIGNED>(sw + 8, sig + 64 * i + 32);
    fe x, y, zt, rx;
    uint8_t s1 = schnorr_phase1(x, y, zt, rx, pkw, mw, sw, tab, gtab, GLoadDev());
    st[j] = s1;
    if (s1 == KGV_ST_PENDING) {
      X[j] = x; Y[j] = y; ZT[j] = zt; RX[j] = rx;
      pre[j] = acc;
      fe_mul(acc, acc, zt);
      any = true;
    }
  }
  if (any) {
    fe inv;
    fe_inv(inv, acc);
#pragma unroll 1
    for (int j = KGV_ITEMS - 1; j >= 0; j--) {
      if (st[j] != KGV_ST_PENDING) continue;
      fe zi;
      fe_mul(zi, inv, pre[j]);
      fe_mul(inv, inv, ZT[j]);
      st[j] = schnorr_phase2(X[j], Y[j], zi, RX[j]);
    }
  }
#pragma unroll 1
  for (int j = 0; j < KGV_ITEMS; j++) {
    size_t i = base + tid + (size_t)j * total;
    if (i < n) status[INDEXED ? index[i] : i] = st[j];
  }
  }
}

struct sc_words { uint32_t v[8]; };

template <bool ALIGNED, bool INDEXED>
__global__ void __launch_bounds__(KGV_BLOCK, KGV_BLOCKS_PER_SM)
k_ecdsa_verify(const uint8_t* __restrict__ pk, const uint8_t* __restrict__ msg, const uint8_t* __restrict__ sig, size_t n_arg,
               uint8_t* __restrict__ status, const uint32_t* __restrict__ gtab, const uint32_t* __restrict__ index, const uint32_t* __restrict__ n_dev) {
  extern __shared__ uint32_t smem[];
  const size_t n = (INDEXED && n_dev) ? (size_t)*n_dev : n_arg;
  const size_t total = (size_t)gridDim.x * KGV_BLOCK;
  const size_t tid = (size_t)blockIdx.x * KGV_BLOCK + threadIdx.x;
  SmemTab tab{smem + threadIdx.x};
  fe QX[KGV_ITEMS], QY[KGV_ITEMS];
  sc_words R_[KGV_ITEMS], S_[KGV_ITEMS], M_[KGV_ITEMS], pre[KGV_ITEMS];
  uint8_t st[KGV_ITEMS];
#pragma unroll 1
  for (size_t base = 0; base < n; base += total * KGV_ITEMS) {
  uint32_t acc[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  bool any = false;
#pragma unroll 1
  for (int j = 0; j < KGV_ITEMS; j++) {
    size_t i = base + tid + (size_t)j * total;
    st[j] = KGV_ST_INVALID;
    if (i >= n) continue;
    if (INDEXED) i = index[i];
    uint32_t pkw[8], mw[8], sw[16];
    const uint8_t* kp = pk + 33 * i;  // 33-byte stride: never word aligned
    uint32_t tag = kp[0];
    load_be32<false>(pkw, kp + 1);
    load_be32<ALIGNED>(mw, msg + 32 * i);
    load_be32<ALIGNED>(sw, sig + 64 * i);
    load_be32<ALIGNED>(sw + 8, sig + 64 * i + 32);
    fe qx, qy;
    uint32_t r[8], s[8], m[8];
    uint8_t s1 = ecdsa_phase1(qx, qy, r, s, m, tag, pkw, mw, sw);
    st[j] = s1;
    if (s1 == KGV_ST_PENDING) {
      QX[j] = qx; QY[j] = qy;
#pragma unroll
      for (int w = 0; w < 8; w++) { R_[j].v[w] = r[w]; S_[j].v[w] = s[w]; M_[j].v[w] = m[w]; pre[j].v[w] = acc[w]; }
      sc_mul(acc, acc, s);
      any = true;
    }
  }
  if (any) {
    uint32_t inv[8];
    sc_inv(inv, acc);
#pragma unroll 1
    for (int j = KGV_ITEMS - 1; j >= 0; j--) {
      if (st[j] != KGV_ST_PENDING) continue;
      uint32_t sn[8];
      sc_mul(sn, inv, pre[j].v);
      sc_mul(inv, inv, S_[j].v);
      st[j] = ecdsa_phase2(QX[j], QY[j], R_[j].v, sn, M_[j].v, tab, gtab, GLoadDev());
    }
  }
#pragma unroll 1
  for (int j = 0; j < KGV_ITEMS; j++) {
    size_t i = base + tid + (size_t)j * total;
    if (i < n) status[INDEXED ? index[i] : i] = st[j];
  }
  }
}

// audit/debug: one signature, every traced intermediate written to dbg[stage*16 ..]
struct DevTrace {
  uint32_t* out;
  __device__ __forceinline__ void operator()(int stage, const uint32_t* w, int n) const {
    for (int i = 0; i < n && i < 16; i++) out[stage * 16 + i] = w[i];
  }
};
__global__ void __launch_bounds__(KGV_BLOCK, KGV_BLOCKS_PER_SM)
k_schnorr_trace(const uint8_t* __restrict__ pk, const uint8_t* __restrict__ msg, const uint8_t* __restrict__ sig,
                uint8_t* __restrict__ status, const uint32_t* __restrict__ gtab, uint32_t* __restrict__ dbg) {
  extern __shared__ uint32_t smem[];
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t pkw[8], mw[8], sw[16];
  load_be32<false>(pkw, pk);
  load_be32<false>(mw, msg);
  load_be32<false>(sw, sig);
  load_be32<false>(sw + 8, sig + 32);
  SmemTab tab{smem + threadIdx.x};
  status[0] = schnorr_verify_core(pkw, mw, sw, tab, gtab, GLoadDev(), DevTrace{dbg});
}

// audit/debug: exercise the arithmetic primitives directly (PTX bodies) on caller-provided operands.
// in: n items x 16 words (a[8], b[8]); out: n items x 16 words.
__global__ void k_selftest(int op, const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[8], b[8], r[16];
#pragma unroll
  for (int k = 0; k < 8; k++) { a[k] = in[i * 16 + k]; b[k] = in[i * 16 + 8 + k]; }
#pragma unroll
  for (int k = 0; k < 16; k++) r[k] = 0;
  fe fa, fb, fr;
#pragma unroll
  for (int k = 0; k < 8; k++) { fa.v[k] = a[k]; fb.v[k] = b[k]; }
  switch (op) {
    case 0: mul_wide(r, a, b); break;
    case 1: sqr_wide(r, a); break;
    case 2: fe_mul(fr, fa, fb); for (int k = 0; k < 8; k++) r[k] = fr.v[k]; break;
    case 3: fe_sqr(fr, fa); for (int k = 0; k < 8; k++) r[k] = fr.v[k]; break;
    case 4: sc_mul(r, a, b); break;
    case 5: sc_sqr(r, a); break;
    case 6: sc_inv(r, a); break;
    case 7: fe_inv(fr, fa); for (int k = 0; k < 8; k++) r[k] = fr.v[k]; break;
    case 8: fe_add(fr, fa, fb); for (int k = 0; k < 8; k++) r[k] = fr.v[k]; break;
    case 9: fe_sub(fr, fa, fb); for (int k = 0; k < 8; k++) r[k] = fr.v[k]; break;
    case 10: { uint32_t t[16]; mul_wide(t, a, b); sc_reduce512(r, t); } break;
    case 11: { uint32_t k1[5], k2[5]; bool n1, n2; glv_split(k1, n1, k2, n2, a); for (int k = 0; k < 5; k++) { r[k] = k1[k]; r[8 + k] = k2[k]; } r[5] = n1; r[13] = n2; } break;
    default: break;
  }
#pragma unroll
  for (int k = 0; k < 16; k++) out[i * 16 + k] = r[k];
}

__global__ void k_status_to_bitmap(const uint8_t* __restrict__ status, size_t n, uint8_t* __restrict__ bitmap) {
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t nbytes = (n + 7) / 8;
  if (b >= nbytes) return;
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    size_t i = 8 * b + j;
    if (i < n && status[i] == KGV_ST_VALID) bits |= 1u << j;
  }
  bitmap[b] = (uint8_t)bits;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      char b_[256];                                                                               \
      snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = b_;                                                                              \
      return KGV_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

static int fail_arg(kgv_ctx* ctx, const char* msg) {
  if (ctx) ctx->err = msg;
  return KGV_ERR_ARG;
}

// 1 = device-accessible pointer, 0 = host pointer
int kgv_ptr_is_device(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Per-call device buffers only ever grow.  The outgrown allocation is NOT freed on the spot: cudaFree synchronises the whole device, which
// stalls every other stream and deadlocks a process that drives several contexts of one device whose kernels wait for each other (the
// peer-exchange wait kernels of kgv_comm.cu); cudaMallocAsync was tried and blocks in the same situation (measured).  Outgrown buffers are
// parked and released by kgv_synchronize / kgv_destroy, i.e. at points where the caller has declared the context idle.  Growth is
// geometric (x1.25), so the parked memory stays below ~4x the live buffer.
int kgv_reserve(kgv_ctx* ctx, uint8_t** buf, size_t* cap, size_t need) {
  if (*cap >= need) return KGV_OK;
  if (*buf) { ctx->parked.push_back(*buf); *buf = nullptr; *cap = 0; }
  size_t want = need + need / 4 + 4096;
  cudaError_t e = cudaMalloc((void**)buf, want);
  if (e != cudaSuccess) {
    // memory pressure: now it is worth a device synchronisation to give the parked buffers back and try again
    (void)cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->aux_stream);
    for (uint8_t* p : ctx->parked) cudaFree(p);
    ctx->parked.clear();
    e = cudaMalloc((void**)buf, want);
  }
  if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e); (void)cudaGetLastError(); *buf = nullptr; return KGV_ERR_NOMEM; }
  *cap = want;
  return KGV_OK;
}

extern "C" int kgv_create(int device, uint32_t flags, kgv_ctx** out) {
  (void)flags;
  if (!out) return KGV_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    (void)cudaGetLastError();
    return KGV_ERR_CUDA;  // no CUDA device: there is no CPU path
  }
  kgv_ctx* ctx = new kgv_ctx();
  ctx->device = device;
  auto body = [&]() -> int {
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CK(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    CK(cudaMalloc((void**)&ctx->gtab, (size_t)2 * 65536 * 16 * sizeof(uint32_t)));
    k_build_gtab<<<(2 * 65536) / 128, 128, 0, ctx->stream>>>(ctx->gtab);
    CK(cudaGetLastError());
    ctx->launches++;
    const int smem = KGV_BLOCK * 128 * (int)sizeof(uint32_t);
    CK(cudaFuncSetAttribute(k_schnorr_verify<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_schnorr_verify<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_schnorr_verify<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_schnorr_verify<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_ecdsa_verify<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_ecdsa_verify<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_ecdsa_verify<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_ecdsa_verify<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_schnorr_verify<true, false>, KGV_BLOCK, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    ctx->resident_blocks = per_sm * sms > 0 ? per_sm * sms : 148 * KGV_BLOCKS_PER_SM;
    CK(cudaStreamSynchronize(ctx->stream));
    return KGV_OK;
  };
  int rc = body();
  if (rc != KGV_OK) {
    fprintf(stderr, "kgv_create: %s\n", ctx->err.c_str());
    delete ctx;
    return rc;
  }
  *out = ctx;
  return KGV_OK;
}

extern "C" void kgv_destroy(kgv_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto& P : ctx->prefetch) if (P.worker.joinable()) P.worker.join();
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->gtab) cudaFree(ctx->gtab);
  for (uint8_t* b : {ctx->d_in, ctx->d_out, ctx->d_batch, ctx->prefetch[0].buf, ctx->prefetch[1].buf, ctx->d_scratch, ctx->d_mu, ctx->d_work, ctx->d_replay})
    if (b) cudaFree(b);
  for (uint8_t* b : ctx->parked) cudaFree(b);
  for (cudaEvent_t e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ev_time) if (e) cudaEventDestroy(e);
  if (ctx->ev_prefetch) cudaEventDestroy(ctx->ev_prefetch);
  for (auto& P : ctx->prefetch) if (P.done) cudaEventDestroy(P.done);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" int kgv_set_stream(kgv_ctx* ctx, void* cuda_stream) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  ctx->stream = (cudaStream_t)cuda_stream;  // NULL is CUDA's default stream, exactly as in cudaStream_t
  return KGV_OK;
}

extern "C" int kgv_reset_stream(kgv_ctx* ctx) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  ctx->stream = ctx->own_stream;
  return KGV_OK;
}

extern "C" int kgv_synchronize(kgv_ctx* ctx) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (!ctx->parked.empty()) {  // the caller declared the context idle: outgrown buffers can go
    CK(cudaStreamSynchronize(ctx->aux_stream));
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
    for (uint8_t* p : ctx->parked) cudaFree(p);
    ctx->parked.clear();
  }
  return KGV_OK;
}

extern "C" const char* kgv_last_error(const kgv_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" uint64_t kgv_launch_count(const kgv_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------
// signature verification entry points
// ---------------------------------------------------------------------------------------------
int kgv_launch_verify(kgv_ctx* ctx, const uint8_t* dpk, const uint8_t* dmsg, const uint8_t* dsig, size_t n, uint8_t* dst, bool ecdsa,
                      cudaStream_t on, bool use_on, const uint32_t* index, const uint32_t* n_dev) {
  if (n == 0) return KGV_OK;
  cudaStream_t st = use_on ? on : ctx->stream;
  const int smem = KGV_BLOCK * 128 * (int)sizeof(uint32_t);
  // one resident wave, persistent; items are strided by the grid size, so a batch smaller than the wave still
  // spreads over every SM (one item per thread) instead of packing KGV_ITEMS items into a quarter of the threads
  size_t want = (n + KGV_BLOCK - 1) / KGV_BLOCK;
  unsigned blocks = (unsigned)(want < (size_t)ctx->resident_blocks ? want : (size_t)ctx->resident_blocks);
  bool aligned = (((uintptr_t)dmsg | (uintptr_t)dsig | (ecdsa ? 0 : (uintptr_t)dpk)) & 31) == 0;
  if (ecdsa) {
    if (index) {
      if (aligned) k_ecdsa_verify<true, true><<<blocks, KGV_BLOCK, smem, st>>>(dpk, dmsg, dsig, n, dst, ctx->gtab, index, n_dev);
      else k_ecdsa_verify<false, true><<<blocks, KGV_BLOCK, smem, st>>>(dpk, dmsg, dsig, n, dst, ctx->gtab, index, n_dev);
    } else {
      if (aligned) k_ecdsa_verify<true, false><<<blocks, KGV_BLOCK, smem, st>>>(dpk, dmsg, dsig, n, dst, ctx->gtab, nullptr, nullptr);
      else k_ecdsa_verify<false, false><<<blocks, KGV_BLOCK, smem, st>>>(dpk, dmsg, dsig, n, dst, ctx->gtab, nullptr, nullptr);
    }
  } else {
    if (index) {
      if (aligned) k_schnorr_verify<true, true><<<blocks, KGV_BLOCK, smem, st>>>(dpk, dmsg, dsig, n, dst, ctx->gtab, index, n_dev);
      else k_schnorr_verify<false, true><<<blocks, KGV_BLOCK, smem, st>>>(dpk, dmsg, dsig, n, dst, ctx->gtab, index, n_dev);
    } else {
      if (aligned) k_schnorr_verify<true, false><<<blocks, KGV_BLOCK, smem, st>>>(dpk, dmsg, dsig, n, dst, ctx->gtab, nullptr, nullptr);
      else k_schnorr_verify<false, false><<<blocks, KGV_BLOCK, smem, st>>>(dpk, dmsg, dsig, n, dst, ctx->gtab, nullptr, nullptr);
    }
  }
  CK(cudaGetLastError());
  ctx->launches++;
  return KGV_OK;
}

static int verify_common(kgv_ctx* ctx, const uint8_t* pk, size_t pk_stride, const uint8_t* msg, const uint8_t* sig, size_t n,
                         uint8_t* status, bool ecdsa) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (n == 0) return KGV_OK;
  if (!pk || !msg || !sig || !status) return fail_arg(ctx, "null buffer");
  CK(cudaSetDevice(ctx->device));
  int dev = kgv_ptr_is_device(pk);
  if (kgv_ptr_is_device(msg) != dev || kgv_ptr_is_device(sig) != dev || kgv_ptr_is_device(status) != dev)
    return fail_arg(ctx, "all buffers of one call must be host pointers or all device pointers");
  const uint8_t *dpk = pk, *dmsg = msg, *dsig = sig;
  uint8_t* dst = status;
  if (!dev) {
    size_t off_msg = (pk_stride * n + 255) & ~(size_t)255;
    size_t off_sig = off_msg + 32 * n;
    int rc = kgv_reserve(ctx, &ctx->d_in, &ctx->d_in_cap, off_sig + 64 * n);
    if (rc) return rc;
    rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, n);
    if (rc) return rc;
    dpk = ctx->d_in; dmsg = ctx->d_in + off_msg; dsig = ctx->d_in + off_sig; dst = ctx->d_out;
    // Large host batches are uploaded in chunks of one full persistent wave (resident threads x KGV_ITEMS signatures) on the
    // side stream while the previous chunk is being verified: only the first chunk's upload is exposed.
    const size_t chunk = (size_t)ctx->resident_blocks * KGV_BLOCK * KGV_ITEMS;
    if (n >= 2 * chunk && (n + chunk - 1) / chunk <= 32) {
      CK(cudaEventRecord(ctx->ev_fork, ctx->stream));          // the staging buffers may still be read by earlier work of this stream
      CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
      size_t c = 0;
      for (size_t a = 0; a < n; a += chunk, c++) {
        const size_t m = n - a < chunk ? n - a : chunk;
        if (!ctx->ev_chunk[c]) CK(cudaEventCreateWithFlags(&ctx->ev_chunk[c], cudaEventDisableTiming));
        CK(cudaMemcpyAsync(ctx->d_in + pk_stride * a, pk + pk_stride * a, pk_stride * m, cudaMemcpyHostToDevice, ctx->aux_stream));
        CK(cudaMemcpyAsync(ctx->d_in + off_msg + 32 * a, msg + 32 * a, 32 * m, cudaMemcpyHostToDevice, ctx->aux_stream));
        CK(cudaMemcpyAsync(ctx->d_in + off_sig + 64 * a, sig + 64 * a, 64 * m, cudaMemcpyHostToDevice, ctx->aux_stream));
        CK(cudaEventRecord(ctx->ev_chunk[c], ctx->aux_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[c], 0));
        int rc2 = kgv_launch_verify(ctx, dpk + pk_stride * a, dmsg + 32 * a, dsig + 64 * a, m, dst + a, ecdsa);
        if (rc2) return rc2;
      }
      CK(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      return KGV_OK;
    }
    CK(cudaMemcpyAsync(ctx->d_in, pk, pk_stride * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_in + off_msg, msg, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_in + off_sig, sig, 64 * n, cudaMemcpyHostToDevice, ctx->stream));
  }
  {
    int rc = kgv_launch_verify(ctx, dpk, dmsg, dsig, n, dst, ecdsa);
    if (rc) return rc;
  }
  if (!dev) {
    CK(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return KGV_OK;
}

extern "C" int kgv_schnorr_verify(kgv_ctx* ctx, const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status) {
  return verify_common(ctx, pk32, 32, msg32, sig64, n, status, false);
}
extern "C" int kgv_ecdsa_verify(kgv_ctx* ctx, const uint8_t* pk33, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status) {
  return verify_common(ctx, pk33, 33, msg32, sig64, n, status, true);
}

extern "C" int kgv_status_to_bitmap(kgv_ctx* ctx, const uint8_t* status, size_t n, uint8_t* bitmap) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (n == 0) return KGV_OK;
  if (!status || !bitmap) return fail_arg(ctx, "null buffer");
  CK(cudaSetDevice(ctx->device));
  int dev = kgv_ptr_is_device(status);
  if (kgv_ptr_is_device(bitmap) != dev) return fail_arg(ctx, "all buffers of one call must be host pointers or all device pointers");
  size_t nbytes = (n + 7) / 8;
  const uint8_t* dsrc = status;
  uint8_t* ddst = bitmap;
  if (!dev) {
    int rc = kgv_reserve(ctx, &ctx->d_in, &ctx->d_in_cap, n);
    if (rc) return rc;
    rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, nbytes);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_in, status, n, cudaMemcpyHostToDevice, ctx->stream));
    dsrc = ctx->d_in; ddst = ctx->d_out;
  }
  k_status_to_bitmap<<<(unsigned)((nbytes + 255) / 256), 256, 0, ctx->stream>>>(dsrc, n, ddst);
  CK(cudaGetLastError());
  ctx->launches++;
  if (!dev) {
    CK(cudaMemcpyAsync(bitmap, ddst, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return KGV_OK;
}

extern "C" int kgv_debug_schnorr_trace(kgv_ctx* ctx, const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, uint32_t* trace_words,
                                       uint8_t* status) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!pk32 || !msg32 || !sig64 || !trace_words || !status) return fail_arg(ctx, "null buffer");
  CK(cudaSetDevice(ctx->device));
  const size_t tw = KGV_TRACE_STAGES * 16 * sizeof(uint32_t);
  int rc = kgv_reserve(ctx, &ctx->d_in, &ctx->d_in_cap, 256);
  if (rc) return rc;
  rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, tw + 256);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_in, pk32, 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_in + 32, msg32, 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_in + 64, sig64, 64, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_out, 0, tw + 256, ctx->stream));
  const int smem = KGV_BLOCK * 128 * (int)sizeof(uint32_t);
  CK(cudaFuncSetAttribute(k_schnorr_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_schnorr_trace<<<1, KGV_BLOCK, smem, ctx->stream>>>(ctx->d_in, ctx->d_in + 32, ctx->d_in + 64, ctx->d_out + tw, ctx->gtab, (uint32_t*)ctx->d_out);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(trace_words, ctx->d_out, tw, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(status, ctx->d_out + tw, 1, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return KGV_OK;
}

extern "C" int kgv_debug_selftest(kgv_ctx* ctx, int op, const uint32_t* in_words, uint32_t* out_words, size_t n) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!in_words || !out_words || n == 0 || n > (1u << 20)) return fail_arg(ctx, "bad selftest arguments");
  CK(cudaSetDevice(ctx->device));
  int rc = kgv_reserve(ctx, &ctx->d_in, &ctx->d_in_cap, n * 64);
  if (rc) return rc;
  rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, n * 64);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_in, in_words, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  k_selftest<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(op, (const uint32_t*)ctx->d_in, (uint32_t*)ctx->d_out, (int)n);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(out_words, ctx->d_out, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return KGV_OK;
}

extern "C" int kgv_gtable_entry(kgv_ctx* ctx, int which, uint32_t v, uint8_t out_xy[64]) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if ((which != 0 && which != 1) || v == 0 || v > 65535 || !out_xy) return fail_arg(ctx, "bad table index");
  CK(cudaSetDevice(ctx->device));
  uint32_t w[16];
  CK(cudaMemcpyAsync(w, ctx->gtab + ((size_t)which * 65536 + v) * 16, sizeof w, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int c = 0; c < 2; c++)
    for (int i = 0; i < 8; i++) {
      uint32_t limb = w[c * 8 + 7 - i];
      out_xy[c * 32 + 4 * i] = (uint8_t)(limb >> 24);
      out_xy[c * 32 + 4 * i + 1] = (uint8_t)(limb >> 16);
      out_xy[c * 32 + 4 * i + 2] = (uint8_t)(limb >> 8);
      out_xy[c * 32 + 4 * i + 3] = (uint8_t)limb;
    }
  return KGV_OK;
}
