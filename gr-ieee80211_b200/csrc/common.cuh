// common.cuh -- internal declarations shared by the kernels and the context (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/c80211b200.h"
#include "lut.h"

#define C8B_SYNC_BUF 240       // lib/sync_impl.h:30
#define C8B_SYNC_RES 111       // lib/sync_impl.h:31
#define C8B_SYM_SHIFT 8        // C8P_SYM_SAMP_SHIFT, lib/cloud80211phy.h:33
#define C8B_DECODE_B_MAX 4095  // lib/decode_impl.h:35
#define C8B_DECODE_T_MAX 32782 // lib/decode_impl.h:36

// Viterbi kernel geometry (k_viterbi.cu)
#define C8B_VIT_CH 150         // trellis steps per chunk: 5 decision groups of 30 (30 = 6 x 5 layout phases)
#define C8B_VIT_WARPS 4        // warps (= frames in flight) per CTA
#define C8B_VIT_TPAD 35040     // uint2 survivor slots per warp: ceil(32782/150) chunks x 160

// kernel launchers (defined next to their kernels); all asynchronous on `st`
void c8b_launch_viterbi(const c8b_lut* d_lut, c8b_frame* d_frames, int nframes, const float* d_llr, int64_t nllr,
                        uint2* d_surv, int nwarps_alloc, uint8_t* d_pdu, int64_t pdu_stride, uint8_t* d_scram,
                        int64_t scram_stride, unsigned* d_counter, int grid, cudaStream_t st);
int c8b_viterbi_max_grid(int num_sm);
