// Pasta curves (Pallas / Vesta: y^2 = x^3 + 5 over two 255-bit primes that are each other's scalar field) for the
// IPA decider -- SURVEY 8f-4: `IpaAs::decide` (snark-verifier/src/pcs/ipa/decider.rs:33-56) checks
// `U == multi_scalar_multiplication(h_coeffs(xi), g)`, the one in-tree caller of the reference's Pippenger
// (util/msm.rs:238-317); the reference's test instantiates it over `halo2_curves::pasta::pallas` (pcs/ipa.rs:407-446).
//
// The field and point code is the BN254 code (field.cuh / g1.cuh templates) over different constants.  Bounds that the
// BN254 comments state for 254-bit moduli also hold here (p < 2^255): a + b < 2p < 2^256; the multiplier's running value
// stays < 2p; the squaring's U + T_hi < 1.5 p.  The fused dot products are NOT instantiated for these fields (their single
// final subtraction needs the 254-bit bound).  Both primes are 1 mod 2^32, so there is no (p+1)/4 square root: points
// cross the ABI uncompressed (64 B affine), as the reference passes `C` values, not encodings, to `decide`.
#pragma once
#include "g1.cuh"

struct PallasBaseParams {
  HD static constexpr u32 mod(int i) {
    constexpr u32 m[8] = {0x00000001u, 0x992d30edu, 0x094cf91bu, 0x224698fcu, 0x00000000u, 0x00000000u, 0x00000000u, 0x40000000u};
    return m[i];
  }
  HD static constexpr u32 one(int i) {
    constexpr u32 m[8] = {0xfffffffdu, 0x34786d38u, 0xe41914adu, 0x992c350bu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu};
    return m[i];
  }
  HD static constexpr u32 r2(int i) {
    constexpr u32 m[8] = {0x0000000fu, 0x8c78ecb3u, 0x8b0de0e7u, 0xd7d30dbdu, 0xc3c95d18u, 0x7797a99bu, 0x7b9cb714u, 0x096d41afu};
    return m[i];
  }
  static constexpr u32 M0 = 0xffffffffu;
};
struct VestaBaseParams {
  HD static constexpr u32 mod(int i) {
    constexpr u32 m[8] = {0x00000001u, 0x8c46eb21u, 0x0994a8ddu, 0x224698fcu, 0x00000000u, 0x00000000u, 0x00000000u, 0x40000000u};
    return m[i];
  }
  HD static constexpr u32 one(int i) {
    constexpr u32 m[8] = {0xfffffffdu, 0x5b2b3e9cu, 0xe3420567u, 0x992c350bu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu};
    return m[i];
  }
  HD static constexpr u32 r2(int i) {
    constexpr u32 m[8] = {0x0000000fu, 0xfc9678ffu, 0x891a16e3u, 0x67bb433du, 0x04ccf590u, 0x7fae2310u, 0x7ccfdaa9u, 0x096d41afu};
    return m[i];
  }
  static constexpr u32 M0 = 0xffffffffu;
};

typedef Fe<PallasBaseParams> PallasFp;  // Pallas base field = Vesta scalar field
typedef Fe<VestaBaseParams> VestaFp;    // Vesta base field  = Pallas scalar field

// Curve traits: base field, scalar field, short-Weierstrass b (a = 0 for all three)
struct CurveBn254 {
  typedef Fq Base;
  typedef Fr Scalar;
  static constexpr u32 B = 3;
};
struct CurvePallas {
  typedef PallasFp Base;
  typedef VestaFp Scalar;
  static constexpr u32 B = 5;
};
struct CurveVesta {
  typedef VestaFp Base;
  typedef PallasFp Scalar;
  static constexpr u32 B = 5;
};

template <class C>
HD bool curve_on_curve(const AffT<typename C::Base>& p) {
  typedef typename C::Base F;
  if (p.is_identity()) return true;
  F b = F::zero(), one = F::one();
  for (u32 i = 0; i < C::B; i++) b = b + one;
  return p.y.sqr() == p.x.sqr() * p.x + b;
}
