// Keccak-256 (Ethereum padding) as an incremental sponge, for the native `EvmTranscript`
// (snark-verifier/src/system/halo2/transcript/evm.rs:152-243): hash = Keccak256(buf || [1 if buf.len() == 32]),
// buf <- hash.  Bytes are absorbed as they are produced (the buffer is never materialised); one state per proof.
#pragma once
#include "field.cuh"

struct KeccakSponge {
  u64 st[25];
  u32 pos;  // bytes absorbed into the current 136-byte block
};

HD u64 rotl64(u64 x, int n) { return (x << n) | (x >> (64 - n)); }

HDN void keccak_f1600(u64* a) {
  const u64 RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
                      0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
                      0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
                      0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
                      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  for (int round = 0; round < 24; round++) {
    u64 c[5], d[5];
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    // rho + pi
    u64 b[25];
    const int rot[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#pragma unroll
    for (int x = 0; x < 5; x++)
#pragma unroll
      for (int y = 0; y < 5; y++) {
        int i = x + 5 * y;
        u64 v = rot[i] ? rotl64(a[i], rot[i]) : a[i];
        b[y + 5 * ((2 * x + 3 * y) % 5)] = v;
      }
    // chi
#pragma unroll
    for (int y = 0; y < 5; y++)
#pragma unroll
      for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ ((~b[(x + 1) % 5 + 5 * y]) & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= RC[round];
  }
}

HD void keccak_reset(KeccakSponge& k) {
#pragma unroll
  for (int i = 0; i < 25; i++) k.st[i] = 0;
  k.pos = 0;
}

HD void keccak_absorb_byte(KeccakSponge& k, uint8_t byte) {
  k.st[k.pos >> 3] ^= (u64)byte << ((k.pos & 7) * 8);
  if (++k.pos == 136) {
    keccak_f1600(k.st);
    k.pos = 0;
  }
}

HD void keccak_absorb(KeccakSponge& k, const uint8_t* p, u32 n) {
  for (u32 i = 0; i < n; i++) keccak_absorb_byte(k, p[i]);
}

// 32 bytes big-endian of a canonical 256-bit value given as 8 little-endian u32 limbs
HD void keccak_absorb_limbs_be(KeccakSponge& k, const u32* v) {
  for (int i = 7; i >= 0; i--) {
    keccak_absorb_byte(k, (uint8_t)(v[i] >> 24));
    keccak_absorb_byte(k, (uint8_t)(v[i] >> 16));
    keccak_absorb_byte(k, (uint8_t)(v[i] >> 8));
    keccak_absorb_byte(k, (uint8_t)v[i]);
  }
}

// Finish the current message: out[32] = Keccak256(absorbed bytes); the sponge is reset.
HD void keccak_finish(KeccakSponge& k, uint8_t* out) {
  k.st[k.pos >> 3] ^= (u64)0x01 << ((k.pos & 7) * 8);
  k.st[16] ^= 0x8000000000000000ull;  // last byte (135) of the 136-byte rate
  keccak_f1600(k.st);
  for (int i = 0; i < 32; i++) out[i] = (uint8_t)(k.st[i >> 3] >> ((i & 7) * 8));
  keccak_reset(k);
}
