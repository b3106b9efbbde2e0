/* c80211b200.h -- C ABI of libc80211b200.so: the B200 (sm_100a) implementation of the
 * gr-ieee80211 20 MHz OFDM receive chain  presiso -> trigger -> sync -> signal -> demod -> decode.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  A gr::block shell
 * (or the ctypes host code in gr-ieee80211_b200/) calls these instead of the reference's scalar C++:
 *
 *   reference interface replaced                          entry point here
 *   ---------------------------------------------------  ---------------------------------------
 *   examples/presiso.grc:35-229 (stock GR hier block)     c8b_presiso
 *   lib/trigger_impl.cc:59-117  trigger::general_work     c8b_trigger        (+ inside c8b_detect)
 *   lib/sync_impl.cc:61-196     sync::general_work        c8b_detect         (trigger+sync+signal)
 *   lib/signal_impl.cc:62-206   signal::general_work      c8b_detect / CFO copy fused in c8b_demod
 *   lib/demod_impl.cc:59-557    demod::general_work       c8b_demod          (header + symbols)
 *   lib/signal2_impl.cc:63-212, lib/demod2_impl.cc:58-806 (2x2) c8b_demod2 / c8b_rx_batch2
 *   lib/decode_impl.cc:60-520   decode::general_work      c8b_decode / c8b_viterbi
 *   whole flowgraph examples/rx.grc:753-767               c8b_rx_batch / c8b_rx_batch_dev
 *   the blocks' general_work(), call by call (all seven)  c8b_blk_create / c8b_blk_work   (gr/lib/rx_blocks_impl.cc)
 *   LUTs of lib/cloud80211phy.cc (c8p.h:151-195)          c8b_lut_blob / c8b_lut_load
 *
 * Conventions: every function returns 0 on success or a negative C8B_ERR_* code (never throws,
 * never aborts; c8b_last_error() gives the text).  The caller owns every buffer it passes.  One
 * c8b_ctx = one CUDA device + one stream + its scratch; a ctx is not re-entrant, different ctxs
 * are independent (thread-per-block safe, like the reference's blocks).  There is NO CPU
 * fallback: without a CUDA device c8b_create fails with C8B_ERR_NO_DEVICE.
 *
 * IQ layout everywhere: interleaved float32 (re,im) = GNU Radio gr_complex = numpy complex64.
 */
#ifndef C80211B200_H
#define C80211B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C8B_ABI_VERSION 1

/* error codes (return values) */
#define C8B_OK 0
#define C8B_ERR_NO_DEVICE (-1)   /* no CUDA device / driver: the library refuses to run       */
#define C8B_ERR_CUDA (-2)        /* a CUDA call failed (text in c8b_last_error)               */
#define C8B_ERR_ARG (-3)         /* bad argument / size over the ctx capacity                 */
#define C8B_ERR_LUT (-4)         /* LUT blob missing, wrong magic/version/size                */
#define C8B_ERR_NOMEM (-5)
#define C8B_ERR_FULL (-6)        /* c8b_stream_push: more frames decided than frames_cap      */

/* per-frame status, in pipeline order (what the reference's blocks do at that point) */
#define C8B_ST_OK 0              /* frame reached decode (PDU count may still be 0: CRC fail)  */
#define C8B_ST_NO_TRIGGER 1      /* no 0x01 trigger in the item   (lib/trigger_impl.cc:101-109) */
#define C8B_ST_SYNC 2            /* LTF autocorr max <= 0.5       (lib/sync_impl.cc:99)         */
#define C8B_ST_LSIG 3            /* every L-SIG check failed      (lib/signal_impl.cc:156-160)  */
#define C8B_ST_TRUNC 4           /* frame runs past the end of the item                         */
#define C8B_ST_FORMAT 5          /* HT/VHT sanity failed -> CLEAN (lib/demod_impl.cc:167,194)   */
#define C8B_ST_DECODE_RANGE 6    /* len>4095 or trellis>32782     (lib/decode_impl.cc:93-97)    */
#define C8B_ST_NDP 7             /* VHT NDP (nSym==0): no PDU                                   */
#define C8B_ST_OVERFLOW 8        /* frame needs more LLR scratch than the ctx sized per frame   */
#define C8B_ST_EMPTY 9           /* unused frame record of an item (max_frames > frames found)  */

/* formats / code rates: same numbering as lib/cloud80211phy.h:35-49 */
#define C8B_F_L 0
#define C8B_F_HT 1
#define C8B_F_VHT 2
#define C8B_CR_12 0
#define C8B_CR_23 1
#define C8B_CR_34 2
#define C8B_CR_56 3

/* One received frame: the union of the tags the reference's blocks attach
 * (sync: rad/snr/rssi lib/sync_impl.cc:124-136; signal: cfo/mcs/len/nsamp lib/signal_impl.cc:135-152;
 * demod: format/mcs/len/cr/ampdu/trellis/total/sssnr lib/demod_impl.cc:224-263) plus bookkeeping. */
typedef struct c8b_frame {
    int32_t status;
    int32_t item;
    int32_t trig_idx;      /* item-relative index of the 0x01 trigger                             */
    int32_t sync_idx;      /* item-relative index of the sync flag (LTF start + 16)               */
    float   rad;           /* CFO compensation step, rad/sample                                   */
    float   snr;
    float   rssi;
    float   cfo_hz;        /* rad * 20e6 / (2 pi)                                                 */
    int32_t l_mcs;         /* L-SIG rate index, length, nsamp                                     */
    int32_t l_len;
    int32_t nsamp;
    int32_t format;        /* C8B_F_*                                                             */
    int32_t mcs;
    int32_t len;
    int32_t cr;            /* C8B_CR_*                                                            */
    int32_t ampdu;
    int32_t nss;
    int32_t nsym;
    int32_t nsymsamp;
    int32_t ncbps;
    int32_t ndbps;
    int32_t trellis;
    int32_t total;         /* nSym * nCBPS soft bits                                              */
    int32_t data_off;      /* sample index (relative to sync_idx+224) of the first DATA symbol    */
    float   sssnr0;
    float   sssnr1;
    int64_t llr_off;       /* offset (floats) of this frame's LLR stream in the LLR arena         */
    int64_t pdu_off;       /* offset (bytes) of this frame's PDU records in the PDU arena         */
    int32_t npdu;          /* PDUs published; record = [fmt][len lo][len hi][MPDU][mcs]           */
    int32_t pdu_bytes;     /* (lib/decode_impl.cc:359-361,414-419,439-441,512-516)                */
} c8b_frame;

typedef struct c8b_cfg {
    int32_t device;          /* CUDA device ordinal                                               */
    int32_t chunk_items;     /* items processed per pipeline pass (scratch is sized for this)     */
    int32_t max_item_len;    /* reserved (scratch is sized on demand from the items of each call)  */
    int32_t max_frames;      /* frame records per item (0 -> 1); a long capture is one item with many */
    int32_t mupos;           /* demod(mupos, mugid) ctor args (lib/demod_impl.cc:28-32)           */
    int32_t mugid;
    int32_t no_overlap;      /* 1: run every kernel on one stream (stage timing); 0: the decode kernel of chunk k overlaps
                                the front end of chunk k+1 on a second stream                            */
    int32_t decode_mode;     /* 0: pick by batch size; 1: one warp per frame pair (k_viterbi, low latency);
                                2: one thread per frame (k_viterbi_tp, throughput)                        */
    int32_t frontend_mode;   /* 0: warp-cooperative detect / header kernels; 1: one thread per item (k_detect, k_header) */
    int32_t mmse;            /* 2x2 equaliser of demod2: 0 = zero forcing (H^H H)^-1 H^H as the reference does
                                (lib/demod2_impl.cc:410-429); 1 = MMSE (H^H H + sigma^2 I)^-1 H^H, unbiased, sigma^2 from sync's
                                snr / rssi tags (north_star "LS/MMSE"; off by default: parity tests run the reference's form) */
    int32_t reserved[4];
} c8b_cfg;

typedef struct c8b_ctx c8b_ctx;

int  c8b_abi_version(void);
/* number of visible CUDA devices (0 or negative error) -- lets hosts fail loudly before create */
int  c8b_device_count(void);
int  c8b_create(const c8b_cfg* cfg, c8b_ctx** out);
void c8b_destroy(c8b_ctx* ctx);
const char* c8b_last_error(const c8b_ctx* ctx);      /* ctx may be NULL: last create error         */
void* c8b_stream(c8b_ctx* ctx);                      /* the ctx's cudaStream_t (for event timing)  */

/* ---- lookup tables (the only thing ranks exchange: one NCCL broadcast at start-up) ------------
 * c8b_lut_blob builds the blob on the host BY FORMULA (interleavers 802.11-2016 17.3.5.7 / 19.3.11.8.3,
 * pilot polarity 17.3.5.10, LTF signs, trellis outputs of g0=133o g1=171o, CRC-32 table); the result
 * equals the reference's transcribed tables c8p.cc:30-31,95-175,1413-1831,1864-1887 (tested). */
size_t c8b_lut_size(void);
int  c8b_lut_blob(void* buf, size_t cap);
int  c8b_lut_load(c8b_ctx* ctx, const void* blob, size_t n);       /* host blob -> device          */
int  c8b_lut_load_dev(c8b_ctx* ctx, const void* d_blob, size_t n); /* device blob (after ncclBcast) */

/* ---- whole chain, batched: items are independent capture segments processed from reset state --
 * iq: complex samples; item i = iq[off[i] .. off[i]+len[i]).  With F = cfg.max_frames (default 1), frames[i*F + k]
 * is the k-th frame accepted by L-SIG in item i, in stream order (record i*F carries the drop status when there is
 * none, unused records are C8B_ST_EMPTY); `frames` holds nitems*F records.  PDU records of frame slot s are written
 * at pdu + s*pdu_stride (at most pdu_stride bytes), `pdu` holds nitems*F*pdu_stride bytes.  c8b_rx_batch takes HOST buffers (pinned or not)
 * and pipelines H2D / kernels / D2H chunk by chunk; c8b_rx_batch_dev takes a DEVICE iq pointer and
 * host off/len/frames/pdu. */
int  c8b_rx_batch(c8b_ctx* ctx, const float* h_iq, const int64_t* off, const int32_t* len, int nitems,
                  c8b_frame* frames, uint8_t* pdu, int64_t pdu_stride);
int  c8b_rx_batch_dev(c8b_ctx* ctx, const float* d_iq, const int64_t* off, const int32_t* len, int nitems,
                      c8b_frame* frames, uint8_t* pdu, int64_t pdu_stride);
/* as c8b_rx_batch with the capture in the radio's wire format: interleaved int16 (I, Q) pairs (UHD "sc16"), half the bytes
 * of fc32 across PCIe.  Each sample is widened on the device to the float the reference's source block (UHD's sc16 -> fc32
 * converter in front of examples/rx.grc) would have produced, x * (1 / 32768), so frames and PDUs equal c8b_rx_batch on the
 * widened capture byte for byte.  off / len in samples. */
int  c8b_rx_batch_sc16(c8b_ctx* ctx, const int16_t* h_iq, const int64_t* off, const int32_t* len, int nitems,
                       c8b_frame* frames, uint8_t* pdu, int64_t pdu_stride);
/* 2x2 receive (examples/rx2.grc:676-692: antenna 0 feeds presiso/trigger/sync, both antennas feed signal2 ->
 * demod2): h_iq0 / h_iq1 are the two antennas' captures with the SAME item offsets/lengths. */
int  c8b_rx_batch2(c8b_ctx* ctx, const float* h_iq0, const float* h_iq1, const int64_t* off, const int32_t* len, int nitems,
                   c8b_frame* frames, uint8_t* pdu, int64_t pdu_stride);
/* Device-pointer entry points run on the ctx stream (created non-blocking): the caller must make sure d_iq is complete
 * (synchronise the producing stream) before calling.
 * as c8b_rx_batch_dev but results stay on the device (d_frames: nitems*F c8b_frame records, d_pdu:
 * nitems*F*pdu_stride bytes, F = cfg.max_frames -- both are cleared / written over their whole extent); nothing is
 * copied back and the call does not synchronise. */
int  c8b_rx_batch_dev_async(c8b_ctx* ctx, const float* d_iq, const int64_t* h_off, const int32_t* h_len, int nitems,
                            c8b_frame* d_frames, uint8_t* d_pdu, int64_t pdu_stride);
/* 2x2 with both antennas resident in device memory (same item table for both), results to host buffers */
int  c8b_rx_batch2_dev(c8b_ctx* ctx, const float* d_iq0, const float* d_iq1, const int64_t* h_off, const int32_t* h_len, int nitems,
                       c8b_frame* frames, uint8_t* pdu, int64_t pdu_stride);
int  c8b_sync(c8b_ctx* ctx);                         /* wait for the ctx stream                    */

/* ---- live stream: what the gr::block shells' general_work calls feed ------------------------------
 * The reference's blocks keep their state between general_work calls (trigger FSM lib/trigger_impl.cc:59-117, sync
 * hold-off lib/sync_impl.cc:73-147, signal's S_COPY position lib/signal_impl.cc:164-192, demod / decode packet
 * state lib/demod_impl.cc:59-342, lib/decode_impl.cc:60-162).  A session keeps a device-resident window of the
 * capture (nant = 1: rx.grc, nant = 2: rx2.grc; the ctx needs max_frames >= 2 = frame records per window pass).
 * c8b_stream_push appends n samples (any n, any split) and returns the frames that became fully decidable:
 * frames[k] (k < *nframes <= frames_cap) with trig_idx / sync_idx relative to absolute stream sample
 * frame_base[k], PDU records at pdu + k*pdu_stride.  flush != 0 ends the stream: what is left is processed with
 * the whole-capture (c8b_rx_batch) semantics, truncated frames included.  The frames and PDUs of a stream are the
 * ones one c8b_rx_batch call over the whole capture (one item) returns, independent of the push sizes.
 * A window that fills up without any decidable point is dropped and counted (c8b_stream_state).  frames_cap must cover what
 * one push can release (a frame is at least 400 samples: n / 400 + max_frames is always enough); C8B_ERR_FULL means frames
 * were lost and the session has to be restarted with c8b_stream_begin. */
int  c8b_stream_begin(c8b_ctx* ctx, int nant, int64_t window_samples /* 0: 4 Mi samples */);
int  c8b_stream_push(c8b_ctx* ctx, const float* h_iq0, const float* h_iq1 /* NULL for nant 1 */, int64_t n, int flush,
                     c8b_frame* frames, int frames_cap, int* nframes, int64_t* frame_base, uint8_t* pdu, int64_t pdu_stride);
int  c8b_stream_state(const c8b_ctx* ctx, int64_t* base, int64_t* fill, int64_t* overruns);

/* ---- the seven receive blocks, one scheduler call at a time ------------------------------------------
 * What the gr::block shells of gr/lib/<block>_impl.cc call from general_work(): same stream signatures, same consume /
 * produce accounting at the stream level, same tags and messages as the reference's blocks
 *   trigger  lib/trigger_impl.cc:59-117   in {float preac}                       out {u8 flags}
 *   sync     lib/sync_impl.cc:61-153      in {u8 trigger, c64 preconj, c64 sig}  out {u8 sync}      tag rad/snr/rssi
 *   signal   lib/signal_impl.cc:62-206    in {u8 sync, c64 sig}                  out {c64}          tag cfo/snr/rssi/seq/mcs/len/nsamp/chan
 *   signal2  lib/signal2_impl.cc:63-212   in {u8 sync, c64 sig0, c64 sig1}       out {c64, c64}     (same tag, on port 0)
 *   demod    lib/demod_impl.cc:59-342     in {c64}                               out {float LLR}    tag cfo/snr/rssi/format/mcs/len/cr/ampdu/trellis/total/sssnr0[/mu2x1chan]
 *   demod2   lib/demod2_impl.cc:58-348    in {c64, c64}                          out {float LLR}    (+ sssnr1)
 *   decode   lib/decode_impl.cc:60-162    in {float LLR}                         message port "out"
 * so examples/rx.grc / rx2.grc connect them unchanged.  The state machines (which call consumes what) run on the host
 * inside the library; every piece of arithmetic -- the trigger FSM over the samples, LTF autocorrelation and CFO, L-SIG
 * FFTs / Viterbi, the CFO-rotated copy, header fields, per-symbol demod, Viterbi decode / CRC -- is a kernel launch on the
 * block's own context.  One block = one c8b_blk = one context: safe under GNU Radio's thread-per-block scheduler. */
#define C8B_BLK_TRIGGER 0
#define C8B_BLK_SYNC 1
#define C8B_BLK_SIGNAL 2
#define C8B_BLK_SIGNAL2 3
#define C8B_BLK_DEMOD 4
#define C8B_BLK_DEMOD2 5
#define C8B_BLK_DECODE 6

/* One stream tag group: the reference attaches ONE pmt dict, split into one tag per key, at one item
 * (lib/sync_impl.cc:124-136, lib/signal_impl.cc:135-152, lib/demod_impl.cc:224-263).  The values travel in `f`
 * (sync: rad snr rssi; signal: cfo_hz snr rssi l_mcs l_len nsamp + seq + vec = chan[64]; demod: cfo_hz snr rssi format mcs
 * len cr ampdu trellis total sssnr0 sssnr1 (+ vec = mu2x1chan[128] for an NDP)). */
typedef struct c8b_tag {
    int32_t port;          /* stream port the tag sits on (always 0 in this chain)                          */
    int32_t idx;           /* item index relative to the first item of this call's buffer on that port      */
    int32_t nvec;          /* complex values in vec: 64 (chan), 128 (mu2x1chan) or 0                        */
    int32_t seq;           /* signal's packet counter (tag "seq", wraps at 1e9)                             */
    c8b_frame f;
    float   vec[256];
} c8b_tag;

typedef struct c8b_blk c8b_blk;
/* cfg->mupos / mugid = demod::make(mupos, mugid); the block creates its own context and loads the tables */
int  c8b_blk_create(const c8b_cfg* cfg, int kind, c8b_blk** out);
void c8b_blk_destroy(c8b_blk* b);
const char* c8b_blk_last_error(const c8b_blk* b);
/* number of input / output stream ports of a block kind (decode: 1 / 0) and their item sizes in bytes */
int  c8b_blk_ports(int kind, int* nin, int* nout, int in_item_bytes[3], int out_item_bytes[2]);
/* forecast(): items required on every input port for noutput output items (1:1 except decode: noutput + 160) */
int  c8b_blk_forecast(int kind, int noutput);
/* general_work(): in[p] / ninput[p] = items available on input port p, out[q] = room for noutput items on output port q;
 * in_tags = the tags on input port 0 inside [0, ninput[0]) with idx relative to in[0].  Returns through *consumed what
 * consume_each() gets, through *produced the return value of general_work, the tags to attach at
 * nitems_written(0) + idx, and for decode the concatenated PDU records published on port "out"
 * ([fmt][len lo][len hi][MPDU][mcs], n = len + 4; NDP report [20][0][4][128 x (re,im) f32], n = 1027).  decode has no
 * output stream: when out_tags is given it receives one record (port = idx = -1) per frame whose Viterbi pass finished,
 * f.npdu telling how many MPDUs passed the CRC (what decode(ifdebug) prints, lib/decode_impl.cc:377-411,456-509). */
int  c8b_blk_work(c8b_blk* b, int noutput, const int* ninput, const void* const* in, void* const* out,
                  const c8b_tag* in_tags, int n_in_tags, int* consumed, int* produced,
                  c8b_tag* out_tags, int out_tag_cap, int* n_out_tags, uint8_t* msg, int msg_cap, int* msg_bytes);

/* per-kernel device time accumulated since the last reset (CUDA events on the ctx stream) */
#define C8B_K_PRESISO 0
#define C8B_K_DETECT 1
#define C8B_K_HEADER 2
#define C8B_K_DEMOD 3
#define C8B_K_VITERBI 4
#define C8B_K_COUNT 5
int  c8b_timing_enable(c8b_ctx* ctx, int on);
int  c8b_timing_read(c8b_ctx* ctx, double ms[C8B_K_COUNT], int64_t launches[C8B_K_COUNT], int reset);

/* ---- transmit waveform synthesiser ---------------------------------------------------
 * The reference's encode -> modulation -> IFFT/CP -> pad chain (lib/encode_impl.cc:130-241, lib/modulation_impl.cc,
 * lib/pad_impl.cc:37-80, lib/cloud80211phy.cc:2594-3161) as its Python twin tools/phy80211.py produces it
 * (genFromMpdu / genFromAmpdu + genFinalSig): 20 MHz, one spatial stream, long GI, legacy MCS 0-7, HT MCS 0-7
 * (non-aggregated MPDU), VHT MCS 0-8 (A-MPDU, group id 0, partial AID 0; psdu_len 0 = NDP).  The samples equal the
 * generator's (float32 rounding apart); they are what the receive entry points above are tested with. */
typedef struct c8b_txframe {
    int32_t format;        /* C8B_F_L / C8B_F_HT / C8B_F_VHT                                         */
    int32_t mcs;           /* L 0-7; HT 0-15 (8-15: two streams); VHT 0-8, + 16 for two space-time streams */
    int64_t psdu_off;      /* byte offset of the MPDU (L, HT) or A-MPDU (VHT) in the PSDU arena       */
    int32_t psdu_len;      /* bytes, <= 4095                                                          */
    float   cfo_hz;        /* carrier offset applied to the frame (genFinalSig cfoHz)                 */
    int64_t out_off;       /* sample index of the frame's first sample in the IQ arena                */
} c8b_txframe;
/* samples of the frame (a multiple of 80), or C8B_ERR_ARG for an unsupported format / mcs / length; no GPU needed */
int  c8b_tx_nsamp(int format, int mcs, int psdu_len);
/* host buffers: the IQ arena (iq_samples complex) is zero-filled, then every frame is written at its out_off.
 * multiplier = genFinalSig's amplitude (the reference recipes use 12), scrambler_seed 1..127 (the generator uses 93) */
int  c8b_tx_batch(c8b_ctx* ctx, const uint8_t* h_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, float multiplier,
                  int scrambler_seed, float* h_iq, int64_t iq_samples);
/* device buffers: only the frames' own samples are written (gaps keep what they hold); asynchronous on the ctx stream */
int  c8b_tx_batch_dev(c8b_ctx* ctx, const uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, float multiplier,
                      int scrambler_seed, float* d_iq, int64_t iq_samples);
/* Two spatial streams on two antennas (lib/encode2_impl.cc, lib/modulation2_impl.cc; generator tools/phy80211.py with nSTS = 2,
 * the recipe of tools/pktGenExample.py:206-217): descriptors with HT mcs 8-15, or VHT mcs 16 + MCS (0-8) for two space-time
 * streams, become stream k on antenna k (direct mapping, cyclic shifts 200 / 400 ns on the second, P-matrix LTFs, power split
 * 1 / sqrt 2; the reference recipes use multiplier 12 sqrt 2); one-stream descriptors in the same batch go to antenna 0 only.
 * Both arenas have iq_samples samples and share the descriptors' out_off.  The one-antenna calls above reject two-stream
 * descriptors. */
int  c8b_tx_batch2(c8b_ctx* ctx, const uint8_t* h_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, float multiplier,
                   int scrambler_seed, float* h_iq0, float* h_iq1, int64_t iq_samples);
int  c8b_tx_batch2_dev(c8b_ctx* ctx, const uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, float multiplier,
                       int scrambler_seed, float* d_iq0, float* d_iq1, int64_t iq_samples);

/* Two-user VHT MU-MIMO transmit waveform (tools/phy80211.py:180-221 genAmpduMu, :571-633 VHT-SIG-B per
 * user, :763-786 data symbols through the spatial mapping; lib/encode2_impl.cc / lib/modulation2_impl.cc carry the same frame in the
 * reference's TX flowgraph, tools/cmu_ap.py drives it).  One space-time stream per user, each with its own A-MPDU (a multiple of
 * 4 bytes) and VHT MCS 0-8, padded to the common symbol count; from the VHT-STF on, antenna t of subcarrier k carries
 * sum_u Q[k][t][u] * X_u[k].  q: nq sets of 64 x 2 x 2 complex floats (re, im), subcarriers in the order -32 .. 31, row = antenna,
 * column = user -- the bfQ list of the generator.  A station receives its frame with cfg.mupos = its user position and
 * cfg.mugid = group_id (lib/demod_impl.cc:238-249). */
typedef struct c8b_txmu {
    int32_t mcs[2];        /* VHT MCS of user 0 / 1                                                     */
    int32_t psdu_len[2];   /* A-MPDU bytes of user 0 / 1 (4 .. 4092, multiple of 4)                     */
    int64_t psdu_off[2];   /* their byte offsets in the PSDU arena                                       */
    int32_t group_id;      /* 1 .. 62                                                                   */
    float   cfo_hz;
    int64_t out_off;       /* first sample of the frame in both antenna arenas                          */
    int64_t q_index;       /* which of the nq matrix sets this frame is mapped with                     */
} c8b_txmu;
int  c8b_tx_mu_nsamp(int mcs0, int len0, int mcs1, int len1);
/* the MAC -> PHY datagrams of the MU demo (host only, no GPU): the two-user frame
 *     [3][mcs0][nss0][len0:2 LE][mcs1][nss1][len1:2 LE][group id][A-MPDU 0][A-MPDU 1]
 * (lib/pktgen_impl.cc:101-113 pktPop, written by tools/phy80211.py:1139-1161 genPktGrDataMu) -> mcs / psdu_len / group_id of *f,
 * *psdu0 / *psdu1 = the A-MPDUs inside pkt; C8B_ERR_ARG for a malformed datagram (pktPop's checks) or one the synthesiser does
 * not make (nss != 1, lengths that are not multiples of 4); and the spatial mapping update
 *     [10][64 x 2 x 2 x (re, im) float32]      (2049 bytes; lib/modulation2_impl.cc:117-121, tools/phy80211.py:1163-1171)
 * -> q_out[512], one matrix set in the layout c8b_tx_mu_batch takes. */
int  c8b_tx_udp_parse_mu(const uint8_t* pkt, int pkt_len, c8b_txmu* f, const uint8_t** psdu0, const uint8_t** psdu1);
int  c8b_tx_udp_parse_bfq(const uint8_t* pkt, int pkt_len, float* q_out);
int  c8b_tx_mu_batch(c8b_ctx* ctx, const uint8_t* h_psdu, int64_t psdu_bytes, const c8b_txmu* frames, int nframes, const float* h_q, int nq,
                     float multiplier, int scrambler_seed, float* h_iq0, float* h_iq1, int64_t iq_samples);
int  c8b_tx_mu_batch_dev(c8b_ctx* ctx, const uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txmu* frames, int nframes, const float* d_q, int nq,
                         float multiplier, int scrambler_seed, float* d_iq0, float* d_iq1, int64_t iq_samples);

/* synthetic traffic for closed-loop runs: fills every frame's PSDU region (device memory) with a random MPDU carrying a valid
 * FCS (tools/mac80211.py:36-47); VHT regions (psdu_len a multiple of 4) get the one-MPDU A-MPDU delimiter of
 * tools/mac80211.py:333-360 in front.  A frame decoded by the receive path returns exactly these bytes. */
int  c8b_tx_random_psdu_dev(c8b_ctx* ctx, uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, uint64_t seed);

/* MAC -> PHY wire format: what the reference's TX chain takes on its UDP socket (lib/pktgen_impl.cc:57-70 msgRead, :96-118
 * pktPop; written by tools/phy80211.py:1126-1137 genPktGrData): one datagram per frame,
 *     [format:1][mcs:1][nss:1][len:2 little endian][len PSDU bytes]      (format 0 L, 1 HT, 2 VHT; VHT PSDU = A-MPDU)
 * c8b_tx_udp_parse (host only, no GPU): checks one datagram the way pktPop does (>= 5 bytes, len <= 4095, the datagram holds
 * len bytes after the header) and fills format / mcs / psdu_len of *f (mcs in the descriptor's coding: VHT + 16 for two
 * streams); *psdu points at the PSDU inside pkt.  Returns the number of spatial streams (1 or 2) or C8B_ERR_ARG for a
 * malformed datagram, the MU format (3, two users per datagram) and stream counts the synthesiser does not make.
 * c8b_tx_from_udp: npkts datagrams (datagram k = pkts[pkt_off[k] .. pkt_off[k] + pkt_len[k])) -> one IQ arena: every
 * accepted one-stream datagram becomes a frame, `gap` zero samples in front of each and after the last
 * (tools/pktGenExample.py gapLen); malformed ones are skipped like the reference does, two-stream ones too (they need the two
 * arenas of c8b_tx_batch2: parse them with c8b_tx_udp_parse).  frames_out (npkts records, may be NULL) gets the
 * descriptors used (out_off = where each frame starts; psdu_len = -1 for a skipped datagram); *iq_used = samples written.
 * Returns the number of frames synthesised or a negative error (C8B_ERR_FULL: iq_cap too small). */
int  c8b_tx_udp_parse(const uint8_t* pkt, int pkt_len, c8b_txframe* f, const uint8_t** psdu);
int  c8b_tx_from_udp(c8b_ctx* ctx, const uint8_t* pkts, const int64_t* pkt_off, const int32_t* pkt_len, int npkts, int gap,
                     float multiplier, int scrambler_seed, float* h_iq, int64_t iq_cap, int64_t* iq_used, c8b_txframe* frames_out);

/* ---- staged entry points (host buffers in/out; used for parity tests and ncu captures) ---------
 * Each mirrors one reference block on whole arrays. */
/* presiso: preac[n] (float), preconj[n] (complex, may be NULL) */
int  c8b_presiso(c8b_ctx* ctx, const float* h_iq, int64_t n, float* h_preac, float* h_preconj);
/* trigger FSM over one array from reset state: out[n] flag bytes (0x01 trigger, 0x02 latch) */
int  c8b_trigger(c8b_ctx* ctx, const float* h_preac, int64_t n, uint8_t* h_out);
/* the trigger scan as the batched path runs it (bitmap words, bulk updates, skipped idle stretches) on a given preac array:
 * events[k] = {index of the 0x01 flag, index latched by the last honoured 0x02 flag (-1: none), restart point before its
 * plateau, 1 if fewer than 240 samples follow (the scan stops there)}.  Only triggers that survive sync's 111-sample
 * hold-off (lib/sync_impl.cc:141-146) are reported; the scan starts at sample `from`.  safe_end: restart point at the end. */
int  c8b_trigger_events(c8b_ctx* ctx, const float* h_preac, int64_t n, int from, int32_t* h_events /* [cap][4] */, int cap, int* count,
                        int* safe_end);
/* detect = presiso + trigger + sync + signal on items; fills status..nsamp of frames[i] and
 * h_chan (64 complex per item: the legacy channel, tag "chan" lib/signal_impl.cc:146-152; may be NULL) */
int  c8b_detect(c8b_ctx* ctx, const float* h_iq, const int64_t* off, const int32_t* len, int nitems,
                c8b_frame* frames, float* h_chan);
/* demod: for frames already detected (status C8B_ST_OK, sync_idx/rad/l_mcs/l_len/nsamp set, chan
 * given): header states + per-symbol demod.  LLRs of frame i at h_llr + frames[i].llr_off, where the
 * call sets llr_off = i*llr_stride. */
int  c8b_demod(c8b_ctx* ctx, const float* h_iq, const int64_t* off, const int32_t* len, int nitems,
               c8b_frame* frames, const float* h_chan, float* h_llr, int64_t llr_stride);
/* demod2: as c8b_demod for the 2-antenna block (1- and 2-stream HT/VHT frames, legacy on antenna 0) */
int  c8b_demod2(c8b_ctx* ctx, const float* h_iq0, const float* h_iq1, const int64_t* off, const int32_t* len, int nitems,
                c8b_frame* frames, const float* h_chan, float* h_llr, int64_t llr_stride);
/* decode: depuncture + Viterbi + descramble + assemble + CRC-32 for frames with cr/trellis/format/
 * len/mcs/ampdu/total/llr_off set.  h_scram (may be NULL): one byte per decoded (still scrambled)
 * bit, frame i at h_scram + i*scram_stride. */
int  c8b_decode(c8b_ctx* ctx, const float* h_llr, int64_t nllr, c8b_frame* frames, int nframes,
                uint8_t* h_pdu, int64_t pdu_stride, uint8_t* h_scram, int64_t scram_stride);

#ifdef __cplusplus
}
#endif
#endif
