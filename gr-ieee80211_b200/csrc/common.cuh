// common.cuh -- internal declarations shared by the kernels and the context (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/c80211b200.h"
#include "lut.h"


// Viterbi kernel geometry (k_viterbi.cu)
#define C8B_VIT_CH 150         // trellis steps per chunk: 5 decision groups of 30 (30 = 6 x 5 layout phases)
#define C8B_VIT_WARPS 4        // warps (= frames in flight) per CTA
#define C8B_VIT_TPAD 35040     // uint2 survivor slots per warp: ceil(32782/150) chunks x 160

#define C8B_VIT_WARP_SLOTS (2 * C8B_VIT_TPAD + 640)   // per warp: two frames' survivors + 1093 group words (uint2 units)

// kernel launchers (defined next to their kernels); all asynchronous on `st`
void c8b_launch_viterbi(const c8b_lut* d_lut, c8b_frame* d_frames, int nframes, const float* d_llr, int64_t nllr,
                        uint2* d_surv, int nwarps_alloc, uint8_t* d_pdu, int64_t pdu_stride, uint8_t* d_scram,
                        int64_t scram_stride, unsigned* d_counter, int grid, cudaStream_t st);
int c8b_viterbi_max_grid(int num_sm);
cudaError_t c8b_viterbi_prepare(void);       // per-device kernel attributes, set on the current device by c8b_create
cudaError_t c8b_viterbi_tp_prepare(void);

void c8b_launch_presiso(const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int maxLen, int64_t outBase,
                        float* preac, float2* preconj, uint32_t* mask, int maskStride, cudaStream_t st);
void c8b_launch_sc16_to_fc32(const short2* in, float2* out, int64_t n, cudaStream_t st);
void c8b_launch_trigger(const float* preac, int64_t n, uint8_t* out, cudaStream_t st);
void c8b_launch_detect(const c8b_lut* lut, const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int itemBase,
                       int maxf, int64_t outBase, const float* preac, const uint32_t* mask, int maskStride, c8b_frame* frames,
                       float2* chan, c8b_scan* scans, cudaStream_t st);
void c8b_launch_header(const c8b_lut* lut, const float2* iq, const int64_t* d_off, int nitems, int maxf, int mupos, c8b_frame* frames,
                       const float2* chan, float2* hinv, int64_t llrStride, float* llr, cudaStream_t st);
void c8b_launch_demod(const c8b_lut* lut, const float2* iq, const int64_t* d_off, int nitems, int maxf, int maxSym, const c8b_frame* frames,
                      const float2* hinv, float* llr, cudaStream_t st);
void c8b_launch_header2(const c8b_lut* lut, const float2* iq0, const float2* iq1, const int64_t* d_off, int nitems, int maxf,
                        int mmse, c8b_frame* frames, const float2* chan, float2* hinv, float2* w2, int64_t llrStride, cudaStream_t st);
void c8b_launch_header2_w(const c8b_lut* lut, const float2* iq0, const float2* iq1, const int64_t* d_off, int nitems, int maxf, int mmse,
                          c8b_frame* frames, const float2* chan, float2* hinv, float2* w2, int64_t llrStride, cudaStream_t st);
void c8b_launch_demod2(const c8b_lut* lut, const float2* iq0, const float2* iq1, const int64_t* d_off, int nitems, int maxf, int maxSym,
                       const c8b_frame* frames, const float2* w2, float* llr, cudaStream_t st);
size_t c8b_viterbi_tp_scratch_bytes(int num_sm, int nframes);
void c8b_launch_viterbi_tp(const c8b_lut* d_lut, c8b_frame* d_frames, int nframes, const float* d_llr, int64_t nllr, void* d_scratch,
                           int num_sm, uint8_t* d_pdu, int64_t pdu_stride, uint8_t* d_scram, int64_t scram_stride, cudaStream_t st);
void c8b_launch_detect_w(const c8b_lut* lut, const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int itemBase,
                         int maxf, int64_t outBase, const float* preac, const uint32_t* mask, int maskStride, c8b_frame* frames,
                         float2* chan, c8b_scan* scans, cudaStream_t st);
void c8b_launch_header_w(const c8b_lut* lut, const float2* iq, const int64_t* d_off, int nitems, int maxf, int mupos, c8b_frame* frames,
                         const float2* chan, float2* hinv, int64_t llrStride, float* llr, cudaStream_t st);
int c8b_viterbi_tp_wave(int num_sm);
void c8b_launch_ndp(c8b_frame* frames, int nframes, const float* llr, int64_t nllr, uint8_t* pdu, int64_t pduStride, cudaStream_t st);
int c8b_tx_geometry_host(int format, int mcs, int len, int* nsym, int* nslots);
size_t c8b_tx_plan_bytes(int nframes);
uint32_t c8b_tx_eof_word(void);
void c8b_tx_scrambler(int seed, uint32_t out[4]);
void c8b_launch_tx(const c8b_lut* lut, const c8b_txframe* d_frames, int nframes, int maxSlots, void* d_plan, const uint8_t* d_psdu,
                   float2* d_out, float2* d_out1, float gain, const uint32_t scr[4], uint32_t eof, cudaStream_t st);
int c8b_tx_nss_host(int format, int mcs);
int c8b_tx_mu_geometry_host(int mcs0, int len0, int mcs1, int len1, int* nsym, int* nslots);
void c8b_launch_tx_mu(const c8b_lut* lut, const c8b_txmu* d_frames, int nframes, int maxSlots, void* d_plan, const uint8_t* d_psdu, const float2* d_q,
                      float2* d_out0, float2* d_out1, float gain, const uint32_t scr[4], uint32_t eof, cudaStream_t st);
void c8b_launch_tx_fill(const c8b_lut* lut, const c8b_txframe* d_frames, int nframes, uint8_t* d_psdu, uint64_t seed, cudaStream_t st);
size_t c8b_detect_multi_scratch(int nitems, int maxLen);   // maxLen: longest item, samples
void c8b_launch_detect_multi(const c8b_lut* lut, const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int itemBase,
                             int maxf, int64_t outBase, const float* preac, const uint32_t* mask, int maskStride, c8b_frame* frames,
                             float2* chan, c8b_scan* scans, void* scratch, int maxLen, cudaStream_t st);
void c8b_launch_trigger_events(const float* d_preac, int n, const int64_t* d_off, const int32_t* d_len, uint32_t* d_mask, const c8b_scan* d_scan,
                               void* scratch, int cap, int32_t* d_out, cudaStream_t st);
// completion flag in mapped pinned memory: enqueue c8b_launch_flag last, then poll with c8b_wait_flag (0 = reached, -1 = stream error)
void c8b_launch_flag(uint32_t* d_flag, uint32_t seq, cudaStream_t st);
int c8b_wait_flag(const volatile uint32_t* h_flag, uint32_t seq, cudaStream_t st);
// one frame per call on a context (ctx.cu; used by blocks.cu): packed staging, one copy each way
int c8b_one_demod_submit(c8b_ctx* ctx, int slot, int nant, const float* iq0, const float* iq1, int n, const c8b_frame* f, const float* chan);
int c8b_one_demod_collect(c8b_ctx* ctx, int slot, int wait, c8b_frame* f, const float** llr_out, int* llr_n);   // 1 done, 0 not yet, < 0 error
#define C8B_ONE_SLOTS 4
int c8b_one_decode_submit(c8b_ctx* ctx, int slot, const c8b_frame* f, const float* llr, int nllr);
int c8b_one_decode_collect(c8b_ctx* ctx, int slot, int wait, c8b_frame* f, const uint8_t** pdu_out);   // 1 done, 0 not yet, < 0 error
// one event of one block per launch (csrc/blocks.cu); res: c8b_blocks::SyncRes / SignalRes, device or mapped host memory
void c8b_launch_one_sync(const float2* d_sig, float conj_re, float conj_im, void* res, cudaStream_t st);
void c8b_launch_one_signal(const c8b_lut* lut, const float2* d_in, const float2* d_in1, float rad, void* res, float2* d_out0, float2* d_out1,
                           int ncopy, cudaStream_t st);
void c8b_launch_one_trigger(void* d_state, const float* d_in, int n, uint8_t* d_out, cudaStream_t st);
// the device copy of the table blob of a context (null until c8b_lut_load); used by blocks.cu
extern "C" const c8b_lut* c8b_ctx_lut(const c8b_ctx* ctx);
