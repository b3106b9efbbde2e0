// lut.h -- the lookup-table blob shared by all kernels (and broadcast between ranks).
// Plain POD, built on the host by formula in lut.cc; lives in device global memory.
// Equivalent reference tables: lib/cloud80211phy.cc:30-31 (SIG demap), :95-140 (LTF signs),
// :160-174 (pilots), :1413-1831 (deinterleave maps), :1864-1887 (trellis), boost::crc_32_type.
#pragma once
#include <stdint.h>

#define C8B_SYNC_BUF 240       // lib/sync_impl.h:30
#define C8B_SYNC_RES 111       // lib/sync_impl.h:31
#define C8B_SYM_SHIFT 8        // C8P_SYM_SAMP_SHIFT, lib/cloud80211phy.h:33
#define C8B_DECODE_B_MAX 4095  // lib/decode_impl.h:35
#define C8B_DECODE_T_MAX 32782 // lib/decode_impl.h:36

#define C8B_LUT_MAGIC 0x4c423843u /* "C8BL" */
#define C8B_LUT_VERSION 6u

struct c8b_lut {
    uint32_t magic, version, bytes, pad0;
    float ltfL[64];            // L-LTF sign per FFT bin (0 on unused bins)
    float ltfNL[64];           // HT/VHT-LTF sign per FFT bin
    float ltfNL22[64];         // second VHT-LTF as seen by user position 1 (pilot bins negated)
    float pilotP[128];         // pilot polarity p_0..p_126 (+ pad)
    float twr[64], twi[64];    // W64^k = exp(-2 pi j k / 64)
    double twdr[32], twdi[32]; // the same in double, k < 32 (per-frame DFTs)
    uint16_t deintL[4][288];   // legacy: nBPSC 1,2,4,6      out[map[i]] = in[i]
    uint16_t deintNL[2][5][416]; // HT/VHT 20 MHz: [iss-1][nBPSCS 1,2,4,6,8]
    int8_t sigDemap[64];       // FFT bin -> deinterleaved SIG llr index (-1: not a data tone)
    uint8_t bmClass[32];       // butterfly k: encoder output (o0*2+o1) of transition 2k --0--> k
    uint8_t binToDataL[64];    // FFT bin -> data index 0..47 in -26..26 order (255: null/pilot)
    uint8_t binToDataNL[64];   // FFT bin -> data index 0..51 in -28..28 order (255: null/pilot)
    uint32_t crc32tab[256];    // reflected 0xEDB88320
    uint32_t crcZ[6][32];      // crcZ[p][i] = CRC register (1<<i) advanced by 64*2^p zero bytes (lane-parallel CRC)
    float pair01[2];           // {0.0f, 1.0f}: read as one 8-byte register pair by k_viterbi_tp (FMUL2 / FFMA2 selectors)
    // k_demod's per-thread views of the tables above (thread j of a symbol owns FFT bins j + 8*k2):
    // demapTab[mode][8*j + k2], mode = 0..3 legacy nBPSC 1,2,4,6; 4..8 HT/VHT nBPSCS 1,2,4,6,8.  0xFFFF: null / pilot bin.
    // Otherwise bits 0-8 = B, bits 9-10 = R: soft bit h*s + c of the bin's data tone (s = max(nBPSC/2, 1), h < nBPSC/s, c < s)
    // lands at B + N_COL*((c + R) mod s) + h*N_COL*s -- the deinterleaver's two permutations in closed form (N_COL 16 / 13).
    alignas(16) uint16_t demapTab[9][64];
    // demapTab2[m][a][8*j + k2], m = nBPSCS 1,2,4,6,8, a = stream of a 2-stream symbol: bits 0-9 = P0, 10-11 = R, 12-13 = beta
    // (k_demod.cu demap2_tone): the stream's deinterleaver (rotated by 22 tones for stream 2) and the stream parser in closed form
    alignas(16) uint16_t demapTab2[5][2][64];
    alignas(16) float tw8[4][8][4]; // tw8[p][j] = W64^(j*2p), W64^(j*(2p+1)) as (re, im, re, im): the 8x8 DFT's twiddles; the 8
                                    // threads of a symbol read one 128-byte line per p
};

void c8b_lut_build(c8b_lut* L);   // host, by formula (lut.cc)

// Scan parameters of one item when it is a window of a live stream (c8b_stream_push): the blocks keep their state across
// general_work calls (lib/trigger_impl.cc:59, sync_impl.cc:61, signal_impl.cc:62); here a window restarts from a point
// where that state is known -- the trigger FSM in its reset state, no sync hold-off pending -- and carries only the
// signal block's consumed-until position.
struct c8b_scan {
    int32_t from;        // in:  first sample the trigger FSM sees (samples before it are presiso history only)
    int32_t pos0;        // in:  signal's consumed-until position, window relative (may be negative)
    int32_t flush;       // in:  1 = end of stream: stalls are final (batch semantics), 0 = stop at the first stall
    int32_t safe;        // out: latest sample s with the FSM reset before s and s >= sync hold-off: everything before s is decided
    int32_t pos;         // out: consumed-until as of `safe`
    int32_t nf;          // out: frame records that belong to the decided part (their trigger precedes `safe`)
    int32_t stalled;     // out: 0 scanned to the end, 1 an event needs samples past the window, 2 frame records used up
    int32_t pad;
};
