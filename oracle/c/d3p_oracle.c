/* Oracle, C restatement of the bit-exact (integer) pieces of the DP-VI update path.  TEST INFRASTRUCTURE ONLY: nothing
 * under d3p_b200/ may link or call this; tests/test_oracle_c.py checks it against the numpy oracle (oracle/chacha.py,
 * oracle/threefry.py, oracle/minibatch.py), RFC 8439 and the Random123 known answers, so that every bit-level parity
 * claim rests on two independent restatements.
 *
 *   chacha20 block / keystream        RFC 8439 section 2.3; reference use: d3p/random/__init__.py:28-32 (jax-chacha-prng)
 *   bits -> uniform [0, 1)            jax.random._uniform bit trick, d3p/random/__init__.py:32,80
 *   threefry2x32 (20 rounds)          Random123; jax.random legacy layouts (split / random_bits), d3p/svi.py:290
 *   feistel sample_indices            d3p/util.py:216-301
 *   poisson_sample_idxs               d3p/minibatch.py:29-39 (stable argsort of the selector bits, reversed)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

#define QR(a, b, c, d)                                  \
  a += b; d ^= a; d = rotl(d, 16); c += d; b ^= c; b = rotl(b, 12); \
  a += b; d ^= a; d = rotl(d, 8);  c += d; b ^= c; b = rotl(b, 7);

/* RFC 8439 2.3: 10 double rounds + feed-forward; word 12 of `in` is the block counter */
void d3po_chacha20_block(const uint32_t in[16], uint32_t out[16]) {
  uint32_t x[16];
  memcpy(x, in, sizeof(x));
  for (int i = 0; i < 10; ++i) {
    QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
    QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}

/* keystream words [0, n) of blocks counter + first_block, counter + first_block + 1, ... (random_bits, 32 bit) */
void d3po_keystream(const uint32_t state[16], uint64_t first_block, uint32_t* out, size_t n) {
  uint32_t st[16], blk[16];
  memcpy(st, state, sizeof(st));
  for (size_t b = 0; b * 16 < n; ++b) {
    st[12] = state[12] + (uint32_t)(first_block + b);
    d3po_chacha20_block(st, blk);
    size_t m = n - b * 16 < 16 ? n - b * 16 : 16;
    memcpy(out + b * 16, blk, m * sizeof(uint32_t));
  }
}

float d3po_bits_to_unit_float(uint32_t bits) {
  union { uint32_t u; float f; } v;
  v.u = (bits >> 9) | 0x3F800000u;
  return v.f - 1.0f;
}

/* Threefry-2x32, 20 rounds, Random123 rotation constants and key-schedule parity */
void d3po_threefry2x32(const uint32_t key[2], uint32_t c0, uint32_t c1, uint32_t out[2]) {
  static const int R[8] = {13, 15, 26, 6, 17, 29, 16, 24};
  const uint32_t ks[3] = {key[0], key[1], key[0] ^ key[1] ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  for (int g = 0; g < 5; ++g) {
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = rotl(x1, R[(g & 1) * 4 + r]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  out[0] = x0; out[1] = x1;
}

/* jax.random legacy random_bits(key, 32, (n,)): counts iota(n) padded to even length, first half paired with second */
void d3po_threefry_random_bits(const uint32_t key[2], uint32_t* out, size_t n) {
  const size_t half = (n + 1) / 2;
  for (size_t j = 0; j < half; ++j) {
    uint32_t y[2];
    const uint32_t c1 = j + half < n ? (uint32_t)(j + half) : 0u;     /* the pad element is a zero count */
    d3po_threefry2x32(key, (uint32_t)j, c1, y);
    out[j] = y[0];
    if (j + half < n) out[j + half] = y[1];
  }
}

/* jax.random.split(key, num) -> [num, 2] = random_bits(key, 32, (2 num,)).reshape(num, 2) */
void d3po_threefry_split(const uint32_t key[2], uint32_t num, uint32_t* out /* 2 num */) {
  d3po_threefry_random_bits(key, out, 2 * (size_t)num);
}

/* d3p/util.py:216-301: ten-round Feistel-like bijection on `bits` bits with cycle walking; rc = 30 round constants
 * (rc[3 j] is odd), positions first_pos .. first_pos + n - 1 */
void d3po_feistel_indices(const uint32_t rc[30], uint32_t capacity, uint32_t first_pos, uint32_t n, uint32_t* out) {
  uint32_t bits = 0;
  while (bits < 32 && ((uint64_t)1 << bits) < (uint64_t)capacity) ++bits;      /* (capacity - 1).bit_length() */
  const uint32_t lo = bits >> 1, up = bits - lo;
  const uint32_t mask_lo = lo >= 32 ? 0xFFFFFFFFu : ((1u << lo) - 1u), mask_up = up >= 32 ? 0xFFFFFFFFu : ((1u << up) - 1u);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t x = first_pos + i;
    do {
      for (int j = 0; j < 10; ++j) {
        const uint32_t xu = x >> lo, xl = x & mask_lo;
        const uint32_t F = ((uint32_t)(xu * rc[3 * j + 1]) >> up) ^ rc[3 * j + 2];
        const uint32_t yu = (F & mask_lo) ^ xl;
        const uint32_t yl = (uint32_t)(xu * rc[3 * j]) & mask_up;
        x = (yu << up) | yl;
      }
    } while (x >= capacity);
    out[i] = x;
  }
}

/* d3p/minibatch.py:29-39: sel = uniform(key, (N,)) <= q; idxs = argsort(sel)[::-1][:cutoff] with a stable sort, i.e. the
 * selected records in descending order followed by the unselected ones in descending order.  Returns the number selected. */
uint32_t d3po_poisson_sample(const uint32_t state[16], float q, uint32_t N, uint32_t cutoff, int32_t* idx_out) {
  uint32_t st[16], blk[16];
  memcpy(st, state, sizeof(st));
  uint8_t* sel = (uint8_t*)malloc(N ? N : 1);
  uint32_t num = 0;
  for (uint32_t b = 0; (uint64_t)b * 16 < N; ++b) {
    st[12] = state[12] + b;
    d3po_chacha20_block(st, blk);
    for (uint32_t i = 0; i < 16 && (uint64_t)b * 16 + i < N; ++i) {
      const uint8_t s = d3po_bits_to_unit_float(blk[i]) <= q;
      sel[(size_t)b * 16 + i] = s;
      num += s;
    }
  }
  uint32_t w = 0;
  for (uint32_t r = N; r-- > 0 && w < cutoff;) if (sel[r]) idx_out[w++] = (int32_t)r;
  for (uint32_t r = N; r-- > 0 && w < cutoff;) if (!sel[r]) idx_out[w++] = (int32_t)r;
  free(sel);
  return num;
}
