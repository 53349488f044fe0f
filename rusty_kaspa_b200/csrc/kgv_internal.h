// kgv_internal.h — shared between the translation units of libkgv.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <mutex>
#include <string>

#ifndef KGV_BLOCK
#define KGV_BLOCK 128         // threads per block of the verification kernels
#endif
#ifndef KGV_BLOCKS_PER_SM
#define KGV_BLOCKS_PER_SM 3   // 3 x 64 KiB of per-thread tables in shared memory
#endif

struct kgv_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  uint32_t* gtab = nullptr;     // [2][65536][16] u32: v*G and v*2^128*G, affine
  uint8_t* d_in = nullptr;      // staging for host-pointer calls
  size_t d_in_cap = 0;
  uint8_t* d_out = nullptr;
  size_t d_out_cap = 0;
  uint64_t launches = 0;
  std::mutex mu;
  std::string err;
};

int kgv_ptr_is_device(const void* p);
int kgv_reserve(kgv_ctx* ctx, uint8_t** buf, size_t* cap, size_t need);
