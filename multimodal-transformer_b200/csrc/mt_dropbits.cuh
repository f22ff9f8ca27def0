// Keep bits of the attention-probability dropout (tcgen05 attention, T <= 128), shared by the stand-alone draw kernel
// (mt_attention_tc.cu) and the LayerNorm forward that draws the NEXT attention's bits under its own memory latency (mt_elementwise.cu).
// Layout: bits[G*B][h][4][128] words, bit i of word (c, q) = key 32 c + i of query q; the draws are exactly the pair-hash draws the
// attention kernels make themselves (pair index ((b_local h + hd) T + q) P2 + (j >> 1), low half -> even key), 16 per word.
#pragma once
#include "mt_common.cuh"

constexpr int MT_BITS_MAXG = 4;
struct MtBitsArgs { int B, T, h, G; DropCfg drop[MT_BITS_MAXG]; };
struct MtBitsJob { MtBitsArgs a; uint32_t* bits; };      // host side: what to draw and where
struct MtBitsKeys { uint32_t key[MT_BITS_MAXG], thr[MT_BITS_MAXG], per_group, n, P2; };

__device__ __forceinline__ MtBitsKeys mt_bits_resolve(const MtBitsArgs& a) {
  MtBitsKeys k;
#pragma unroll
  for (int g = 0; g < MT_BITS_MAXG; ++g) {
    const DropCfg d = mt_drop_resolve(a.drop[g < a.G ? g : 0]);
    k.key[g] = d.key; k.thr[g] = d.thresh != 0u ? (d.thresh >> 16) : 0x10000u;      // 16-bit threshold; 0x10000: dropout off, every draw is below it
  }
  k.per_group = (uint32_t)(a.B * a.h) * 512u;               // words of one group
  k.n = k.per_group * (uint32_t)a.G;
  k.P2 = (uint32_t)(a.T + 1) >> 1;
  return k;
}

// word idx of the launch (idx < k.n)
__device__ __forceinline__ uint32_t mt_bits_word(const MtBitsArgs& a, const MtBitsKeys& k, uint32_t idx) {
  const uint32_t q = idx & 127u, c = (idx >> 7) & 3u;
  const uint32_t grp = idx / k.per_group;
  const uint32_t blh = (idx - grp * k.per_group) >> 9;                   // group-local narrative * h + head
  const uint32_t key = grp == 0 ? k.key[0] : (grp == 1 ? k.key[1] : (grp == 2 ? k.key[2] : k.key[3]));
  const uint32_t t16 = grp == 0 ? k.thr[0] : (grp == 1 ? k.thr[1] : (grp == 2 ? k.thr[2] : k.thr[3]));
  uint32_t word = 0xffffffffu;
  if (t16 != 0x10000u && q < (uint32_t)a.T) {
    word = 0u;
    const uint32_t pbase = (blh * (uint32_t)a.T + q) * k.P2 + c * 16u;
    const uint32_t thi = t16 << 16;
    // mt_mix32 with its tail folded into the compares: with y the value before the last xor-shift, the draw is b = y ^ (y >> 16), so
    // its high half is y's (hi >= t16 <=> y >= t16 << 16) and its low half is y_lo ^ y_hi = the high half of y ^ (y << 16).
    // P2 % 16 == 0 (T = 128): pbase has four zero low bits and (pbase + i) ^ key = (pbase ^ key) ^ i.
    auto draw16 = [&](auto seed_of) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t y = seed_of(i);
        y ^= y >> 16;
        y *= 0x7FEB352Du;
        y ^= y >> 15;
        y *= 0x846CA68Bu;
        const uint32_t z = y ^ (y << 16);
        word |= (z >= thi ? (1u << (2 * i)) : 0u) | (y >= thi ? (2u << (2 * i)) : 0u);
      }
    };
    if ((k.P2 & 15u) == 0u) {
      const uint32_t pk = pbase ^ key;
      draw16([&](int i) { return pk ^ (uint32_t)i; });
    } else {
      draw16([&](int i) { return (pbase + (uint32_t)i) ^ key; });
    }
  }
  return word;
}
