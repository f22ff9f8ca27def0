// Argument blocks of the MFN recurrence kernels, shared by the FFMA kernels (mt_mfn.cu: fp32 mode, any size) and the
// tensor-core kernels (mt_mfn_mma.cu: bf16 mode, weights resident in registers).
#pragma once
#include "mt_ops.cuh"

struct LstmArgs {
  int B, T, n_mods;
  long long sb, st;
  int H[MT_MAX_MODS], hoff[MT_MAX_MODS];
  int Hs, MEM;
  const void* w_hh[MT_MAX_MODS];     // WT [4H][H]
  const float* b_hh[MT_MAX_MODS];
  float* gates;                      // [M,4Hs]
  float* cstar;                      // [M,2Hs]
  void* cstar_op;                    // ST [M,2Hs] (null when it aliases cstar)
  void* last_op;                     // ST [M,Hs+MEM]
  void* hprev_op;                    // ST [M,Hs] (training)
  float* h_last; float* c_last;
  int training;
  const float* dlast;                // backward: [M,Hs+MEM]
  const float* dcstar;               //           [M,2Hs]
  void* dz_op;                       //           ST [M,4Hs]
  int dbg;                           // timing experiments (mt_tune key 8): bit 0 no gate stash, 1 no state stores, 2 no feed, 3 no mma
};

struct MemArgs {
  int B, T;
  long long sb, st;
  int Hs, MEM, G;
  const void* g1_fc1_w; const void* g2_fc1_w;      // WT [G][2Hs+MEM]
  const void* g1_fc2_w; const void* g2_fc2_w;      // WT [MEM][G]
  const float* g1_fc2_b; const float* g2_fc2_b;
  const float* gpre;       // [M,2G]
  const float* chat;       // [M,MEM]
  void* gh_op;             // ST [M,2G]
  float* gm;               // [M,2MEM]
  void* memprev_op;        // ST [M,MEM]
  void* last_op;           // ST [M,Hs+MEM]
  float* mem_last;
  DropCfg drop_g1, drop_g2;
  int training;
  const float* dlast;      // backward: [M,Hs+MEM]
  void* dzg_op;            //           ST [M,2MEM]
  void* dzchat_op;         //           ST [M,MEM]
  void* dgh_op;            //           ST [M,2G]
  int dbg;
};


// ---- tensor-core recurrences (mt_mfn_mma.cu), bf16 operands ---------------------------------------------------------
bool mt_mfn_mma_lstm_supported(const LstmArgs& a);
bool mt_mfn_mma_mem_supported(const MemArgs& a);
int mt_mfn_mma_lstm_fwd(const LstmArgs& a, cudaStream_t st);
int mt_mfn_mma_lstm_bwd(const LstmArgs& a, cudaStream_t st);
int mt_mfn_mma_mem_fwd(const MemArgs& a, cudaStream_t st);
int mt_mfn_mma_mem_bwd(const MemArgs& a, cudaStream_t st);
