/* ORACLE — CPU restatement of the reference's MuHash (test infrastructure only, see ok_oracle.h).
 *   crypto/muhash/src/lib.rs:59-166      MuHash {numerator, denominator}, add/remove/combine/finalize/serialize
 *   crypto/muhash/src/u3072.rs:22-193    arithmetic modulo 2^3072 - 1103717
 *   consensus/core/src/muhash.rs:16-60   which bytes of a transaction / UTXO become elements
 * The element expansion uses rand_chacha::ChaCha20Rng (un-vendored dependency, rand_chacha 0.3: djb ChaCha20,
 * 64-bit block counter from 0, stream id 0, keystream words little-endian): restated here from the published
 * algorithm and pinned by the reference's own known answers (lib.rs:17-21 EMPTY_MUHASH, :189-238 TEST_VECTORS,
 * :290-298 test_new_pre_computed, :301-327 test_serialize) in tests/test_oracle_muhash.py.
 * Values are kept canonical (in [0, p)) at all times; the reference's transient "overflown" representations
 * (u3072.rs:49-57) are unobservable through finalize()/serialize(). */
#include <string.h>
#include "ok_oracle.h"

#define L 48
#define PRIME_DIFF 1103717ull
typedef unsigned __int128 u128;

static void chacha20_block(const uint32_t key[8], uint64_t counter, uint8_t out[64]) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                    (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
  uint32_t x[16];
  memcpy(x, s, sizeof x);
#define ROTL(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define QR(a, b, c, d) a += b; d ^= a; d = ROTL(d, 16); c += d; b ^= c; b = ROTL(b, 12); a += b; d ^= a; d = ROTL(d, 8); c += d; b ^= c; b = ROTL(b, 7);
  for (int r = 0; r < 10; r++) {
    QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
    QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 16; i++) {
    uint32_t v = x[i] + s[i];
    out[4 * i] = (uint8_t)v; out[4 * i + 1] = (uint8_t)(v >> 8); out[4 * i + 2] = (uint8_t)(v >> 16); out[4 * i + 3] = (uint8_t)(v >> 24);
  }
}

/* r >= p ?  (u3072.rs:49-57) */
static int u_is_overflow(const uint64_t* a) {
  if (a[0] <= UINT64_MAX - PRIME_DIFF) return 0;
  for (int i = 1; i < L; i++) if (a[i] != UINT64_MAX) return 0;
  return 1;
}
/* a -= p, for a in [p, 2^3072): a + PRIME_DIFF mod 2^3072 (u3072.rs:78-88) */
static void u_full_reduce(uint64_t* a) {
  u128 c = PRIME_DIFF;
  for (int i = 0; i < L; i++) { c += a[i]; a[i] = (uint64_t)c; c >>= 64; }
}
static void u_canon(uint64_t* a) { if (u_is_overflow(a)) u_full_reduce(a); }

/* r = a * b mod p, inputs < 2^3072, output canonical */
static void u_mul(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t t[2 * L + 1];
  memset(t, 0, sizeof t);
  for (int i = 0; i < L; i++) {
    u128 c = 0;
    for (int j = 0; j < L; j++) { c += (u128)a[i] * b[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
    t[i + L] = (uint64_t)c;
  }
  /* fold: lo + hi * PRIME_DIFF, twice */
  uint64_t x[L + 1];
  u128 c = 0;
  for (int i = 0; i < L; i++) { c += (u128)t[L + i] * PRIME_DIFF + t[i]; x[i] = (uint64_t)c; c >>= 64; }
  x[L] = (uint64_t)c; /* < 2^21 */
  c = (u128)x[L] * PRIME_DIFF;
  for (int i = 0; i < L; i++) { c += x[i]; r[i] = (uint64_t)c; c >>= 64; }
  if (c) { /* wrapped once more: value = r + 2^3072 == r + PRIME_DIFF */
    c = PRIME_DIFF;
    for (int i = 0; i < L; i++) { c += r[i]; r[i] = (uint64_t)c; c >>= 64; }
  }
  u_canon(r);
}
static int u_is_zero(const uint64_t* a) { for (int i = 0; i < L; i++) if (a[i]) return 0; return 1; }

/* r = a^-1 mod p by Fermat (p prime), 0 -> 0 (u3072.rs:157-173: "0/x is 0") */
static void u_inverse(uint64_t* r, const uint64_t* a) {
  uint64_t base[L], e[L], acc[L], tmp[L];
  memcpy(base, a, sizeof base);
  u_canon(base);
  if (u_is_zero(base)) { memset(r, 0, 8 * L); return; }
  /* e = p - 2 = 2^3072 - 1103719 */
  for (int i = 0; i < L; i++) e[i] = UINT64_MAX;
  e[0] = UINT64_MAX - (PRIME_DIFF + 2) + 1;
  memset(acc, 0, sizeof acc); acc[0] = 1;
  for (int bit = 3071; bit >= 0; bit--) {
    u_mul(tmp, acc, acc); memcpy(acc, tmp, sizeof acc);
    if ((e[bit >> 6] >> (bit & 63)) & 1) { u_mul(tmp, acc, base); memcpy(acc, tmp, sizeof acc); }
  }
  memcpy(r, acc, 8 * L);
}

static void u_from_le(uint64_t* r, const uint8_t* b) { for (int i = 0; i < L; i++) { uint64_t w = 0; for (int k = 7; k >= 0; k--) w = (w << 8) | b[8 * i + k]; r[i] = w; } }
static void u_to_le(uint8_t* b, const uint64_t* a) { for (int i = 0; i < L; i++) for (int k = 0; k < 8; k++) b[8 * i + k] = (uint8_t)(a[i] >> (8 * k)); }

/* lib.rs:152-165: 32-byte element hash -> 3072-bit element (first 384 keystream bytes, little-endian) */
void ok_muhash_expand(const uint8_t hash32[32], uint8_t out384[384]) {
  uint32_t key[8];
  for (int i = 0; i < 8; i++) key[i] = (uint32_t)hash32[4 * i] | ((uint32_t)hash32[4 * i + 1] << 8) | ((uint32_t)hash32[4 * i + 2] << 16) | ((uint32_t)hash32[4 * i + 3] << 24);
  for (uint64_t blk = 0; blk < 6; blk++) chacha20_block(key, blk, out384 + 64 * blk);
}

void ok_muhash_init(ok_muhash* m) { memset(m, 0, sizeof *m); m->num[0] = 1; m->den[0] = 1; }

static void mul_into(uint64_t* field, const uint8_t elem384[384]) {
  uint64_t e[L], t[L];
  u_from_le(e, elem384);
  u_mul(t, field, e);
  memcpy(field, t, sizeof t);
}
/* lib.rs:61-74 */
void ok_muhash_add_element(ok_muhash* m, const void* data, size_t n) {
  uint8_t h[32], e[384];
  ok_blake2b_keyed("MuHashElement", data, n, h);
  ok_muhash_expand(h, e);
  mul_into(m->num, e);
}
void ok_muhash_remove_element(ok_muhash* m, const void* data, size_t n) {
  uint8_t h[32], e[384];
  ok_blake2b_keyed("MuHashElement", data, n, h);
  ok_muhash_expand(h, e);
  mul_into(m->den, e);
}
/* lib.rs:91-96 */
void ok_muhash_combine(ok_muhash* m, const ok_muhash* o) {
  uint64_t t[L];
  u_mul(t, m->num, o->num); memcpy(m->num, t, sizeof t);
  u_mul(t, m->den, o->den); memcpy(m->den, t, sizeof t);
}
/* lib.rs:105-115: normalize + to_le_bytes */
void ok_muhash_serialize(ok_muhash* m, uint8_t out384[384]) {
  uint64_t inv[L], t[L];
  u_inverse(inv, m->den);
  u_mul(t, m->num, inv);
  memcpy(m->num, t, sizeof t);
  memset(m->den, 0, sizeof m->den); m->den[0] = 1;
  u_to_le(out384, m->num);
}
/* lib.rs:98-102 */
void ok_muhash_finalize(ok_muhash* m, uint8_t out32[32]) {
  uint8_t ser[384];
  ok_muhash_serialize(m, ser);
  ok_blake2b_keyed("MuHashFinalize", ser, 384, out32);
}
/* lib.rs:117-121: 0 ok, -1 OverflowError */
int ok_muhash_deserialize(ok_muhash* m, const uint8_t in384[384]) {
  ok_muhash_init(m);
  u_from_le(m->num, in384);
  return u_is_overflow(m->num) ? -1 : 0;
}
void ok_muhash_raw(const ok_muhash* m, uint8_t num384[384], uint8_t den384[384]) { u_to_le(num384, m->num); u_to_le(den384, m->den); }

/* consensus/core/src/muhash.rs:47-59 write_utxo */
static size_t write_utxo(uint8_t* buf, const uint8_t txid[32], uint32_t index, uint64_t daa, uint64_t amount, int is_coinbase, uint16_t spk_version,
                         const uint8_t* script, uint32_t script_len) {
  size_t o = 0;
  memcpy(buf, txid, 32); o = 32;
  for (int i = 0; i < 4; i++) buf[o++] = (uint8_t)(index >> (8 * i));
  for (int i = 0; i < 8; i++) buf[o++] = (uint8_t)(daa >> (8 * i));
  for (int i = 0; i < 8; i++) buf[o++] = (uint8_t)(amount >> (8 * i));
  buf[o++] = is_coinbase ? 1 : 0;
  buf[o++] = (uint8_t)spk_version; buf[o++] = (uint8_t)(spk_version >> 8);
  uint64_t l = script_len;
  for (int i = 0; i < 8; i++) buf[o++] = (uint8_t)(l >> (8 * i));
  memcpy(buf + o, script, script_len); o += script_len;
  return o;
}
/* muhash.rs:28-33 add_utxo */
void ok_muhash_add_utxo(ok_muhash* m, const uint8_t key36[36], const ok_utxo_entry* e, const uint8_t* bytes) {
  static __thread uint8_t buf[36 + 27 + 65536];
  uint32_t index = (uint32_t)key36[32] | ((uint32_t)key36[33] << 8) | ((uint32_t)key36[34] << 16) | ((uint32_t)key36[35] << 24);
  size_t n = write_utxo(buf, key36, index, e->block_daa_score, e->amount, e->is_coinbase, e->spk_version, bytes + e->script_off, e->script_len);
  ok_muhash_add_element(m, buf, n);
}
/* muhash.rs:16-27 add_transaction: every populated input is removed, every output added with (tx id, i), block_daa_score, is_coinbase */
void ok_muhash_add_transaction(ok_muhash* m, const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint64_t block_daa_score) {
  static __thread uint8_t buf[36 + 27 + 65536];
  const ok_tx* t = &b->txs[tx];
  uint8_t id[32];
  ok_tx_id(b, tx, id);
  static const uint8_t coinbase_subnet[20] = {1};
  int is_coinbase = memcmp(t->subnetwork_id, coinbase_subnet, 20) == 0;
  for (uint32_t i = 0; i < t->n_inputs; i++) {
    const ok_input* in = &b->inputs[t->first_input + i];
    const ok_utxo_entry* e = &entries[t->first_input + i];
    size_t n = write_utxo(buf, in->prev_txid, in->prev_index, e->block_daa_score, e->amount, e->is_coinbase, e->spk_version, b->bytes + e->script_off, e->script_len);
    ok_muhash_remove_element(m, buf, n);
  }
  for (uint32_t i = 0; i < t->n_outputs; i++) {
    const ok_output* o = &b->outputs[t->first_output + i];
    size_t n = write_utxo(buf, id, i, block_daa_score, o->value, is_coinbase, o->spk_version, b->bytes + o->script_off, o->script_len);
    ok_muhash_add_element(m, buf, n);
  }
}
/* utxo_validation.rs:282-309: MuHash::from_transaction of every accepted tx, combined */
void ok_muhash_accepted(ok_muhash* m, const ok_batch* b, const ok_utxo_entry* entries, const uint8_t* accept, uint64_t pov_daa_score) {
  ok_muhash_init(m);
  for (size_t i = 0; i < b->n_txs; i++) if (accept[i]) ok_muhash_add_transaction(m, b, entries, i, pov_daa_score);
}
