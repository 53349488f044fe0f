// kgv_muhash.cu — K8: MuHash product trees and the element-level entry points of include/kgv.h.
//
// Data flow: one thread builds one 3072-bit element (keyed BLAKE2b -> ChaCha20 keystream) straight into the level-0
// array of its tree; each tree level halves the array with one u3072 multiplication per thread (kgv_u3072.cuh) until
// one value is left; the last kernel writes the canonical residue.  The two trees of a MuHash (numerator = added
// elements, denominator = removed elements, crypto/muhash/src/lib.rs:32-35) run on two streams.
// Algorithmic traffic per multiplication: 2 x 384 B read, 384 B written (+ 1.5 KB scratch row written and re-read);
// the work is 9 216 IMAD.WIDE per multiplication: integer-issue bound like the signature kernels.
#include "kgv_internal.h"
#include "kgv_muhash.cuh"

#include <cstdio>

using namespace kgv;

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

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline unsigned nblk(size_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

// one level of a product tree: out[t] = in[2t] * in[2t+1] (the odd element out is copied)
__global__ void __launch_bounds__(128) k_u3072_tree_level(const uint32_t* __restrict__ in, size_t n_in, uint32_t* __restrict__ out, uint32_t* __restrict__ wide) {
  const size_t n_out = (n_in + 1) / 2;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  if (2 * t + 1 < n_in) {
    u3072_mul_mod(out, n_out, t, wide, n_out, t, in, n_in, 2 * t, in, n_in, 2 * t + 1);
  } else {
    for (int i = 0; i < KGV_U3072_BLOCKS; i++) {
      uint32_t r[8];
      u3072_load_block(r, in, n_in, 2 * t, i);
      u3072_store_block(out, n_out, t, i, r);
    }
  }
}
// the same level with one multiplication per 16-lane group (kgv_u3072.cuh, cooperative form): ~10x shorter per-level
// latency, used while a level has too few multiplications to fill the machine with one thread each
#define KGV_COOP_GROUPS 8  // groups (multiplications) per 128-thread block
__global__ void __launch_bounds__(128) k_u3072_tree_level_coop(const uint32_t* __restrict__ in, size_t n_in, uint32_t* __restrict__ out) {
  __shared__ U3072Coop sm[KGV_COOP_GROUPS];
  const size_t n_out = (n_in + 1) / 2;
  const int grp = threadIdx.x >> 4, lane = threadIdx.x & 15;
  const size_t t = (size_t)blockIdx.x * KGV_COOP_GROUPS + grp;
  const bool have = t < n_out;
  const bool mul = have && 2 * t + 1 < n_in;
  u3072_coop_mul_mod(sm[grp], lane, mul, out, n_out, t, in, n_in, 2 * t, in, n_in, 2 * t + 1);
  if (have && !mul && lane < KGV_U3072_BLOCKS) {
    uint32_t r[8];
    u3072_load_block(r, in, n_in, 2 * t, lane);
    u3072_store_block(out, n_out, t, lane, r);
  }
}

// canonical residue of a single value (stride `s`, element 0) as 384 little-endian bytes; n == 0: the value one
__global__ void k_u3072_emit(const uint32_t* __restrict__ a, size_t s, int is_empty, uint32_t* __restrict__ out96) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (is_empty) {
    out96[0] = 1;
    for (int i = 1; i < 96; i++) out96[i] = 0;
    return;
  }
  uint32_t r[96];
  u3072_canonical(r, a, s, 0);
  for (int i = 0; i < 96; i++) out96[i] = r[i];
}

// raw elements: element i = data[offsets[i] .. offsets[i+1]); remove[i] != 0 puts it in the denominator tree.
// Both trees have n slots; the slot of the other tree holds the identity.
__global__ void __launch_bounds__(128) k_muhash_raw_elements(const uint8_t* __restrict__ data, const uint64_t* __restrict__ offsets, const uint8_t* __restrict__ remove,
                                                             size_t n, uint32_t* __restrict__ e_den, uint32_t* __restrict__ e_num) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Blake2b h;
  b2b_init_muhash_element(h);
  const uint64_t a = offsets[i], b = offsets[i + 1];
  for (uint64_t k = a; k < b; k++) b2b_byte(h, data[k]);
  uint64_t d[4];
  b2b_final(h, d);
  const bool rem = remove && remove[i];
  muhash_expand_store(rem ? e_den : e_num, n, i, d);
  u3072_store_one(rem ? e_num : e_den, n, i);
}

// ---------------------------------------------------------------------------------------------
// tree driver
// ---------------------------------------------------------------------------------------------
struct MuTreeMem { uint32_t *lvl0, *lvl1, *wide; };
static size_t tree_bytes(size_t n) { size_t h = (n + 1) / 2; return al256(n * 384) + al256(h * 384) + al256(h * 768) + 512; }
static MuTreeMem tree_mem(uint8_t* base, size_t n) {
  size_t h = (n + 1) / 2;
  MuTreeMem m;
  m.lvl0 = (uint32_t*)base;
  m.lvl1 = (uint32_t*)(base + al256(n * 384));
  m.wide = (uint32_t*)(base + al256(n * 384) + al256(h * 384));
  return m;
}
static uint32_t* tree_out(uint8_t* base, size_t n) { size_t h = (n + 1) / 2; return (uint32_t*)(base + al256(n * 384) + al256(h * 384) + al256(h * 768)); }

int kgv_mu_reserve(kgv_ctx* ctx, size_t n_den, size_t n_num, uint32_t** e_den, uint32_t** e_num) {
  size_t need = tree_bytes(n_den) + tree_bytes(n_num);
  int rc = kgv_reserve(ctx, &ctx->d_mu, &ctx->d_mu_cap, need);
  if (rc) return rc;
  *e_den = tree_mem(ctx->d_mu, n_den).lvl0;
  *e_num = tree_mem(ctx->d_mu + tree_bytes(n_den), n_num).lvl0;
  return KGV_OK;
}

static int reduce_one(kgv_ctx* ctx, uint8_t* base, size_t n, cudaStream_t st) {
  MuTreeMem m = tree_mem(base, n);
  uint32_t* cur = m.lvl0;
  uint32_t* nxt = m.lvl1;
  size_t k = n;
  while (k > 1) {
    size_t h = (k + 1) / 2;
    if (h <= 8192) k_u3072_tree_level_coop<<<nblk(h, KGV_COOP_GROUPS), 128, 0, st>>>(cur, k, nxt);
    else k_u3072_tree_level<<<nblk(h, 128), 128, 0, st>>>(cur, k, nxt, m.wide);
    CK(cudaGetLastError());
    ctx->launches++;
    uint32_t* t = cur; cur = nxt; nxt = t;
    k = h;
  }
  k_u3072_emit<<<1, 32, 0, st>>>(cur, 1, n == 0 ? 1 : 0, tree_out(base, n));
  CK(cudaGetLastError());
  ctx->launches++;
  return KGV_OK;
}

int kgv_mu_reduce(kgv_ctx* ctx, size_t n_den, size_t n_num, uint8_t* out_num384, uint8_t* out_den384) {
  cudaStream_t st = ctx->stream, sx = ctx->aux_stream;
  uint8_t* b_den = ctx->d_mu;
  uint8_t* b_num = ctx->d_mu + tree_bytes(n_den);
  CK(cudaEventRecord(ctx->ev_fork, st));
  CK(cudaStreamWaitEvent(sx, ctx->ev_fork, 0));
  int rc = reduce_one(ctx, b_den, n_den, st);
  if (rc) return rc;
  rc = reduce_one(ctx, b_num, n_num, sx);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev_join, sx));
  CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
  const bool dev = kgv_ptr_is_device(out_num384);
  CK(cudaMemcpyAsync(out_den384, tree_out(b_den, n_den), 384, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out_num384, tree_out(b_num, n_num), 384, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  if (!dev) CK(cudaStreamSynchronize(st));
  return KGV_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI: element level
// ---------------------------------------------------------------------------------------------
extern "C" int kgv_muhash_elements(kgv_ctx* ctx, const uint8_t* data, const uint64_t* offsets, const uint8_t* remove, size_t n, uint8_t* numerator384,
                                   uint8_t* denominator384) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!numerator384 || !denominator384 || (n && (!offsets || !data))) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  if (kgv_ptr_is_device(numerator384) != kgv_ptr_is_device(denominator384)) { ctx->err = "outputs must both be host or both be device pointers"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  uint32_t *e_den = nullptr, *e_num = nullptr;
  int rc = kgv_mu_reserve(ctx, n, n, &e_den, &e_num);
  if (rc) return rc;
  if (n) {
    const uint8_t* ddata = data;
    const uint64_t* doff = offsets;
    const uint8_t* drem = remove;
    if (!kgv_ptr_is_device(offsets)) {
      uint64_t total = offsets[n];
      size_t o_off = al256(total + 8), o_rem = al256(o_off + (n + 1) * 8);
      rc = kgv_reserve(ctx, &ctx->d_in, &ctx->d_in_cap, al256(o_rem + n));
      if (rc) return rc;
      CK(cudaMemcpyAsync(ctx->d_in, data, total, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(ctx->d_in + o_off, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
      if (remove) CK(cudaMemcpyAsync(ctx->d_in + o_rem, remove, n, cudaMemcpyHostToDevice, ctx->stream));
      ddata = ctx->d_in; doff = (const uint64_t*)(ctx->d_in + o_off); drem = remove ? ctx->d_in + o_rem : nullptr;
    }
    k_muhash_raw_elements<<<nblk(n, 128), 128, 0, ctx->stream>>>(ddata, doff, drem, n, e_den, e_num);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  return kgv_mu_reduce(ctx, n, n, numerator384, denominator384);
}

// a <- a * b for both fields (crypto/muhash/src/lib.rs:91-96 combine); all four are 384-byte little-endian residues
__global__ void k_muhash_combine(uint32_t* __restrict__ w) {  // w: [a_num | a_den | b_num | b_den] contiguous words; one warp: group 0 numerators, group 1 denominators
  __shared__ U3072Coop sm[2];
  const int grp = threadIdx.x >> 4, lane = threadIdx.x & 15;
  uint32_t* a = w + 96 * grp;
  const uint32_t* b = w + 96 * (2 + grp);
  u3072_coop_mul_mod(sm[grp], lane, true, a, 1, 0, a, 1, 0, b, 1, 0);
  if (lane == 0) {
    uint32_t r[96];
    u3072_canonical(r, a, 1, 0);
    for (int i = 0; i < 96; i++) a[i] = r[i];
  }
}
extern "C" int kgv_muhash_combine(kgv_ctx* ctx, uint8_t* num_a, uint8_t* den_a, const uint8_t* num_b, const uint8_t* den_b) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!num_a || !den_a || !num_b || !den_b) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  int rc = kgv_reserve(ctx, &ctx->d_mu, &ctx->d_mu_cap, 4096);
  if (rc) return rc;
  uint8_t* w = ctx->d_mu;
  const cudaMemcpyKind in = kgv_ptr_is_device(num_a) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind out = kgv_ptr_is_device(num_a) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CK(cudaMemcpyAsync(w, num_a, 384, in, ctx->stream));
  CK(cudaMemcpyAsync(w + 384, den_a, 384, in, ctx->stream));
  CK(cudaMemcpyAsync(w + 768, num_b, 384, in, ctx->stream));
  CK(cudaMemcpyAsync(w + 1152, den_b, 384, in, ctx->stream));
  k_muhash_combine<<<1, 32, 0, ctx->stream>>>((uint32_t*)w);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(num_a, w, 384, out, ctx->stream));
  CK(cudaMemcpyAsync(den_a, w + 384, 384, out, ctx->stream));
  if (out == cudaMemcpyDeviceToHost) CK(cudaStreamSynchronize(ctx->stream));
  return KGV_OK;
}

// finalize (lib.rs:98-115): serialized = numerator / denominator mod p (canonical), hash = BLAKE2b-256 keyed "MuHashFinalize".
// The inverse is denominator^(p-2): p - 2 = (2^3051 - 1) * 2^21 + 993433, i.e. 3 072 squarings and ~30 multiplications,
// strictly sequential.  OFF the data-parallel path (the reference finalizes once per chain block in sequential code):
// one 16-lane group multiplying cooperatively, a few ms.  w: [num | den | cur | saved] contiguous words; out: 96 words serialized + 8 words hash.
// one real function for the many call sites of the exponentiation ladder (contiguous operands, stride 1)
static __device__ __noinline__ void coop_mul_contig(U3072Coop* sm, int lane, bool act, uint32_t* r, const uint32_t* a, const uint32_t* b) {
  u3072_coop_mul_mod(*sm, lane, act, r, 1, 0, a, 1, 0, b, 1, 0);
}
__global__ void k_muhash_finalize(uint32_t* __restrict__ w, uint32_t* __restrict__ out) {
  __shared__ U3072Coop sm;
  const int lane = threadIdx.x & 15;
  const bool act = threadIdx.x < 16;  // one 16-lane group does the arithmetic; the other half of the warp only keeps the warp syncs company
  uint32_t *num = w, *den = w + 96, *cur = w + 192, *saved = w + 288;
  auto mul = [&](uint32_t* r, const uint32_t* a, const uint32_t* b) { coop_mul_contig(&sm, lane, act, r, a, b); };
  auto copy = [&](uint32_t* d, const uint32_t* s_) {
    if (act) for (int i = lane; i < 96; i += 16) d[i] = s_[i];
    __syncwarp();
  };
  // cur = den^(2^k - 1) with k following the bits of 3051 = 0b101111101011 from the top
  copy(cur, den);
  int k = 1;
  for (int bit = 10; bit >= 0; bit--) {
    copy(saved, cur);
    for (int q = 0; q < k; q++) mul(cur, cur, cur);
    mul(cur, cur, saved);
    k *= 2;
    if ((3051 >> bit) & 1) { mul(cur, cur, cur); mul(cur, cur, den); k += 1; }
  }
  for (int bit = 20; bit >= 0; bit--) {
    mul(cur, cur, cur);
    if ((993433u >> bit) & 1u) mul(cur, cur, den);
  }
  mul(num, num, cur);
  if (threadIdx.x == 0) {
    uint32_t r[96];
    u3072_canonical(r, num, 1, 0);
    Blake2b h;
    b2b_init_muhash_finalize(h);
    for (int i = 0; i < 96; i++) { out[i] = r[i]; b2b_u32(h, r[i]); }
    uint64_t d[4];
    b2b_final(h, d);
    for (int i = 0; i < 4; i++) { out[96 + 2 * i] = (uint32_t)d[i]; out[96 + 2 * i + 1] = (uint32_t)(d[i] >> 32); }
  }
}
extern "C" int kgv_muhash_finalize(kgv_ctx* ctx, const uint8_t* numerator384, const uint8_t* denominator384, uint8_t* serialized384, uint8_t* hash32) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!numerator384 || !denominator384 || !hash32) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  int rc = kgv_reserve(ctx, &ctx->d_mu, &ctx->d_mu_cap, 8192);
  if (rc) return rc;
  uint8_t* w = ctx->d_mu;
  uint8_t* o = w + 4096;
  const bool dev = kgv_ptr_is_device(numerator384);
  CK(cudaMemcpyAsync(w, numerator384, 384, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(w + 384, denominator384, 384, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  k_muhash_finalize<<<1, 32, 0, ctx->stream>>>((uint32_t*)w, (uint32_t*)o);
  CK(cudaGetLastError());
  ctx->launches++;
  const bool odev = kgv_ptr_is_device(hash32);
  if (serialized384) CK(cudaMemcpyAsync(serialized384, o, 384, odev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(hash32, o + 384, 32, odev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
  if (!odev) CK(cudaStreamSynchronize(ctx->stream));
  return KGV_OK;
}
