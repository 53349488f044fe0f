// kgv_muhash.cuh — MuHash element construction on the device (one element per thread).
//   consensus/core/src/muhash.rs:47-59   write_utxo: which bytes of an (outpoint, entry) pair are hashed
//   crypto/muhash/src/lib.rs:152-166     element = ChaCha20Rng(seed = BLAKE2b-256 keyed "MuHashElement")[0..384] as LE integer
//   crypto/muhash/src/lib.rs:98-102      finalize = BLAKE2b-256 keyed "MuHashFinalize" of the 384 serialized bytes
#pragma once
#include "kgv_blake2b.cuh"
#include "kgv_u3072.cuh"

namespace kgv {

// keyed BLAKE2b-256 for the two MuHash domains: the (<= 16 byte) key is given as two little-endian words and placed
// directly into the key block.  Deliberately NO local byte array for the key: nvcc's stack colouring was caught
// overlapping such an array with another live local array (tools/repro/).
KGV_HD void b2b_init_keyed_words(Blake2b& h, uint64_t k0, uint64_t k1, uint32_t keylen) {
  b2b_init(h, B2B_UNKEYED);  // zeroes the block
  h.h[0] = kB2bIV[0] ^ (0x01010000ull ^ ((uint64_t)keylen << 8) ^ 32ull);
  h.m[0] = k0;
  h.m[1] = k1;
  h.fill = 128;  // a full key block is pending: compressed as a middle block when data follows, as the last block otherwise
  h.t = 128;
  h.fresh = false;
}
KGV_HD void b2b_init_muhash_element(Blake2b& h) { b2b_init_keyed_words(h, 0x6C4568736148754Dull, 0x000000746E656D65ull, 13); }   // "MuHashElement"
KGV_HD void b2b_init_muhash_finalize(Blake2b& h) { b2b_init_keyed_words(h, 0x694668736148754Dull, 0x0000657A696C616Eull, 14); }  // "MuHashFinalize"

// digest of write_utxo(outpoint, entry): txid given as 8 little-endian u32 words
KGV_HD void muhash_utxo_digest(uint64_t* d4, const uint32_t* txid8, uint32_t index, uint64_t block_daa_score, uint64_t amount, bool is_coinbase,
                               uint32_t spk_version, const uint8_t* script, uint32_t script_len) {
  Blake2b h;
  b2b_init_muhash_element(h);
#pragma unroll
  for (int w = 0; w < 8; w++) b2b_u32(h, txid8[w]);
  b2b_u32(h, index);
  b2b_u64(h, block_daa_score);
  b2b_u64(h, amount);
  b2b_u8(h, is_coinbase ? 1u : 0u);
  b2b_u16(h, spk_version);
  b2b_u64(h, script_len);
  b2b_bytes(h, script, script_len);
  b2b_final(h, d4);
}

}  // namespace kgv
