// Device image of a compiled protocol (compiler.h -> CompiledProtocol).
#pragma once
#include "g1.cuh"
#include "tape.cuh"

struct PointSched {
  u32 byte_offset;
  u32 reg_x, reg_y;
};

struct MsmTermDev {
  int32_t fixed;  // 1: fixed_bases[base] (g = index n_pre), 0: proof point ordinal `base`
  int32_t base;
  int32_t slot;   // out_scalar slot, -1 => scalar is the constant 1
};

struct ProtocolDevice {
  int mos = 0;
  bool verify_valid = true;
  std::string invalid_reason;
  TapeOp* d_ops = nullptr;
  u32 n_ops = 0, read_ops_end = 0;
  uint16_t* d_aux = nullptr;
  Fr* d_consts = nullptr;
  u32 n_regs = 0, n_instances = 0, n_challenges = 0, n_scalar_slots = 0, proof_len = 0;
  int n_perm = 0;
  size_t n_fr_mul = 0;
  std::vector<u32> num_instance;
  std::vector<PointSched> points;
  PointSched* d_sched = nullptr;
  MsmTermDev *d_lhs = nullptr, *d_rhs = nullptr;
  u32 n_lhs = 0, n_rhs = 0;
  G1Affine* d_fixed = nullptr;  // preprocessed..., then g at index n_pre
  u32 n_pre = 0;
  int dk = -1;                  // deciding key whose g1 is baked in as `svk.g`
};
