// blocks.cu -- C ABI c8b_blk_*: the seven receive blocks of the reference, one scheduler call at a time
// (include/c80211b200.h).  The state machines are the host templates of blocks.h; this file is their sm_100a backend:
//   k_one_trigger_w : the trigger FSM (lib/trigger_impl.cc:59-117) continued from the state the previous call left in device
//                     memory, a bitmap word (32 samples) per step in closed form (k_frontend_w.cu)
//   k_one_sync_w    : ltf_autoCorrelation + ltf_cfo of one trigger (lib/sync_impl.cc:155-196), the batch path's warp routine
//   k_one_signal_w  : L-SIG of one sync flag (lib/signal_impl.cc:108-162) AND the S_COPY samples of the same call (:164-192)
//                     in one launch
//   k_blk_cfo_copy  : S_COPY of the later calls of a frame, one thread per sample
//   demod / demod2 / decode run the batch path's kernels on ONE frame (k_header_w / k_header2_w, k_demod / k_demod2,
//   k_viterbi) through c8b_one_demod / c8b_one_decode (ctx.cu): packed staging, one copy each way.
// What a scheduler call costs here is driver calls and one host <-> device round trip, so every op is: inputs -> one
// pinned staging buffer -> ONE asynchronous H2D copy -> kernel(s) -> small results written straight into mapped host
// memory (a posted write, no copy call) / bulk results in ONE D2H copy -> ONE synchronise.
// No CPU path: c8b_blk_create fails without a CUDA device like c8b_create does.
#include <new>
#include <string>

#include "common.cuh"
#include "phy_serial.cuh"
#include "blocks.h"

namespace {

using c8b::cf;
using c8b_blocks::SignalRes;
using c8b_blocks::SyncRes;

__global__ void __launch_bounds__(256)
k_blk_cfo_copy(const float2* __restrict__ in0, const float2* __restrict__ in1, float2* __restrict__ out0, float2* __restrict__ out1,
               int n, int copied, float rad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cf w = c8b::cis(c8b::fmul((float)(copied + i + 224), rad));          // lib/signal_impl.cc:172-173
    const cf a = c8b::cmul(c8b::mk(in0[i].x, in0[i].y), w);
    out0[i] = make_float2(a.re, a.im);
    if (in1) {
        const cf b = c8b::cmul(c8b::mk(in1[i].x, in1[i].y), w);
        out1[i] = make_float2(b.re, b.im);
    }
}

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct c8b_blk {
    int kind = 0, device = 0;
    c8b_ctx* ctx = nullptr;
    cudaStream_t st = nullptr;
    const c8b_lut* lut = nullptr;
    std::string err;
    // device scratch of the small kernels; h_pin = pinned + mapped host staging ([results 4 KB][bulk]), d_res its device alias
    c8b::TrigState* d_trig = nullptr;
    Buf d_a, d_b, d_c, d_d;
    uint8_t* h_pin = nullptr;
    uint8_t* d_res = nullptr;
    size_t pinCap = 0;
    uint32_t seq = 0;                                             // completion flag at h_pin + 3584 (see c8b_launch_flag)
    // block state (blocks.h)
    c8b_blocks::SyncState sy;
    c8b_blocks::SignalState sg;
    c8b_blocks::DemodState dm;
    c8b_blocks::DecodeState dc;

    int fail(const char* what, cudaError_t e)
    {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return C8B_ERR_CUDA;
    }
    int grow(Buf& b, size_t bytes)
    {
        if (b.cap >= bytes) return C8B_OK;
        if (b.p) cudaFree(b.p);
        b.p = nullptr; b.cap = 0;
        const size_t want = bytes + bytes / 2 + 4096;
        const cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) return fail("cudaMalloc", e);
        b.cap = want;
        return C8B_OK;
    }
#define BK(call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(#call, e_); } while (0)

    int pin(size_t bulk)                                          // pinned staging: 4 KB of mapped results + `bulk` bytes
    {
        if (pinCap >= bulk + 4096) return C8B_OK;
        if (h_pin) { cudaStreamSynchronize(st); cudaFreeHost(h_pin); h_pin = nullptr; pinCap = 0; }
        const size_t want = 4096 + bulk + bulk / 2 + 65536;
        cudaError_t e = cudaHostAlloc((void**)&h_pin, want, cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&d_res, h_pin, 0);
        if (e != cudaSuccess) return fail("cudaHostAlloc(mapped)", e);
        pinCap = want;
        return C8B_OK;
    }

    int finish()                                                  // completion flag last in the stream, then poll it from the host
    {
        seq++;
        c8b_launch_flag(reinterpret_cast<uint32_t*>(d_res + 3584), seq, st);
        if (c8b_wait_flag(reinterpret_cast<volatile uint32_t*>(h_pin + 3584), seq, st) != 0) {
            const cudaError_t e = cudaStreamSynchronize(st);
            return fail("block op", e != cudaSuccess ? e : cudaGetLastError());
        }
        return C8B_OK;
    }

    // ---- Ops backend of blocks.h ----
    int trigger(const float* in, int n, uint8_t* out)
    {
        int r;
        if ((r = grow(d_a, (size_t)n * sizeof(float))) || (r = grow(d_b, (size_t)n)) || (r = pin((size_t)n * 5))) return r;
        float* hin = reinterpret_cast<float*>(h_pin + 4096);
        uint8_t* hout = h_pin + 4096 + (size_t)n * sizeof(float);
        memcpy(hin, in, (size_t)n * sizeof(float));
        BK(cudaMemcpyAsync(d_a.p, hin, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
        c8b_launch_one_trigger(d_trig, (const float*)d_a.p, n, (uint8_t*)d_b.p, st);
        BK(cudaGetLastError());
        BK(cudaMemcpyAsync(hout, d_b.p, (size_t)n, cudaMemcpyDeviceToHost, st));
        if ((r = finish())) return r;
        memcpy(out, hout, (size_t)n);
        return C8B_OK;
    }
    int sync_at(const float* sig, const float conj[2], SyncRes* res)
    {
        int r;
        if ((r = grow(d_a, 240 * sizeof(float2))) || (r = pin(240 * sizeof(float2)))) return r;
        memcpy(h_pin + 4096, sig, 240 * sizeof(float2));
        BK(cudaMemcpyAsync(d_a.p, h_pin + 4096, 240 * sizeof(float2), cudaMemcpyHostToDevice, st));
        c8b_launch_one_sync((const float2*)d_a.p, conj[0], conj[1], d_res, st);     // result: a posted write into mapped host memory
        BK(cudaGetLastError());
        if ((r = finish())) return r;
        memcpy(res, h_pin, sizeof(*res));
        return C8B_OK;
    }
    // L-SIG at in0 (>= 224 samples) and, in the same round trip, ncopy samples behind it CFO-corrected (both antennas of
    // signal2); *rot0 / *rot1 point into the pinned staging and stay valid until the next op
    int signal_at(const float* in0, const float* in1, int ncopy, float rad, SignalRes* res, const float** rot0, const float** rot1)
    {
        const size_t nin = (size_t)(224 + ncopy) * sizeof(float2), nout = (size_t)ncopy * sizeof(float2);
        const size_t a1 = (nin + 255) & ~(size_t)255, o1 = (nout + 255) & ~(size_t)255;
        int r;
        if ((r = grow(d_a, 2 * a1)) || (r = grow(d_b, 2 * o1 + 256)) || (r = pin(2 * a1 + 2 * o1))) return r;
        uint8_t* hin = h_pin + 4096;
        uint8_t* hout = hin + 2 * a1;
        memcpy(hin, in0, nin);
        if (in1) memcpy(hin + a1, in1, nin);
        BK(cudaMemcpyAsync(d_a.p, hin, in1 ? a1 + nin : nin, cudaMemcpyHostToDevice, st));
        c8b_launch_one_signal(lut, (const float2*)d_a.p, in1 ? (const float2*)((uint8_t*)d_a.p + a1) : nullptr, rad, d_res, (float2*)d_b.p,
                              in1 ? (float2*)((uint8_t*)d_b.p + o1) : nullptr, ncopy, st);
        BK(cudaGetLastError());
        if (ncopy > 0) BK(cudaMemcpyAsync(hout, d_b.p, in1 ? o1 + nout : nout, cudaMemcpyDeviceToHost, st));
        if ((r = finish())) return r;
        memcpy(res, h_pin, sizeof(*res));
        *rot0 = ncopy > 0 ? reinterpret_cast<const float*>(hout) : nullptr;
        *rot1 = (ncopy > 0 && in1) ? reinterpret_cast<const float*>(hout + o1) : nullptr;
        return C8B_OK;
    }
    int cfo_copy(const float* in0, const float* in1, float* out0, float* out1, int n, int copied, float rad)
    {
        const size_t bytes = (size_t)n * sizeof(float2), al = (bytes + 255) & ~(size_t)255;
        int r;
        if ((r = grow(d_a, 2 * al)) || (r = grow(d_b, 2 * al)) || (r = pin(4 * al))) return r;
        uint8_t* hin = h_pin + 4096;
        uint8_t* hout = hin + 2 * al;
        memcpy(hin, in0, bytes);
        if (in1) memcpy(hin + al, in1, bytes);
        BK(cudaMemcpyAsync(d_a.p, hin, in1 ? al + bytes : bytes, cudaMemcpyHostToDevice, st));
        k_blk_cfo_copy<<<(n + 255) / 256, 256, 0, st>>>((const float2*)d_a.p, in1 ? (const float2*)((uint8_t*)d_a.p + al) : nullptr, (float2*)d_b.p,
                                                        in1 ? (float2*)((uint8_t*)d_b.p + al) : nullptr, n, copied, rad);
        BK(cudaGetLastError());
        BK(cudaMemcpyAsync(hout, d_b.p, in1 ? al + bytes : bytes, cudaMemcpyDeviceToHost, st));
        if ((r = finish())) return r;
        memcpy(out0, hout, bytes);
        if (in1) memcpy(out1, hout + al, bytes);
        return C8B_OK;
    }
    // demod: two staging slots, frames collected in submission order
    int mq_head = 0, mq_n = 0;
    int demod_submit(int nant, const float* iq0, const float* iq1, int n, const c8b_frame* f, const float* chan)
    {
        if (mq_n >= 2) return C8B_ERR_FULL;
        const int rc = c8b_one_demod_submit(ctx, (mq_head + mq_n) & 1, nant, iq0, iq1, n, f, chan);
        if (rc) { err = c8b_last_error(ctx); return rc; }
        mq_n++;
        return C8B_OK;
    }
    int demod_collect(bool wait, c8b_frame* f, const float** soft, int* nsoft)
    {
        if (mq_n == 0) return 0;
        const int rc = c8b_one_demod_collect(ctx, mq_head, wait ? 1 : 0, f, soft, nsoft);
        if (rc < 0) { err = c8b_last_error(ctx); return rc; }
        if (rc == 1) { mq_head ^= 1; mq_n--; }
        return rc;
    }
    // decode: frames go through the ctx's staging slots round robin; collected in submission order
    int dq_head = 0, dq_n = 0;
    int decode_submit(const c8b_frame* f, const float* llr, int nllr)
    {
        if (dq_n >= C8B_ONE_SLOTS) return C8B_ERR_FULL;          // (decode_work publishes before it submits: at most one gathering + 3 in flight)
        const int slot = (dq_head + dq_n) % C8B_ONE_SLOTS;
        const int rc = c8b_one_decode_submit(ctx, slot, f, llr, nllr);
        if (rc) { err = c8b_last_error(ctx); return rc; }
        dq_n++;
        return C8B_OK;
    }
    int decode_collect(bool wait, c8b_frame* f, const uint8_t** pdu)
    {
        if (dq_n == 0) return 0;
        const int rc = c8b_one_decode_collect(ctx, dq_head, wait ? 1 : 0, f, pdu);
        if (rc < 0) { err = c8b_last_error(ctx); return rc; }
        if (rc == 1) { dq_head = (dq_head + 1) % C8B_ONE_SLOTS; dq_n--; }
        return rc;
    }
#undef BK
};

static std::string g_blkErr;

extern "C" {

int c8b_blk_ports(int kind, int* nin, int* nout, int in_item_bytes[3], int out_item_bytes[2])
{
    static const int NIN[7] = { 1, 3, 2, 3, 1, 2, 1 }, NOUT[7] = { 1, 1, 1, 2, 1, 1, 0 };
    static const int INB[7][3] = { { 4, 0, 0 }, { 1, 8, 8 }, { 1, 8, 0 }, { 1, 8, 8 }, { 8, 0, 0 }, { 8, 8, 0 }, { 4, 0, 0 } };
    static const int OUTB[7][2] = { { 1, 0 }, { 1, 0 }, { 8, 0 }, { 8, 8 }, { 4, 0 }, { 4, 0 }, { 0, 0 } };
    if (kind < 0 || kind > C8B_BLK_DECODE) return C8B_ERR_ARG;
    if (nin) *nin = NIN[kind];
    if (nout) *nout = NOUT[kind];
    for (int k = 0; k < 3; k++) if (in_item_bytes) in_item_bytes[k] = INB[kind][k];
    for (int k = 0; k < 2; k++) if (out_item_bytes) out_item_bytes[k] = OUTB[kind][k];
    return C8B_OK;
}

int c8b_blk_forecast(int kind, int noutput)
{
    if (kind < 0 || kind > C8B_BLK_DECODE || noutput < 0) return C8B_ERR_ARG;
    return kind == C8B_BLK_DECODE ? noutput + 160 : noutput;     // lib/decode_impl.cc:55-58; 1:1 everywhere else
}

const char* c8b_blk_last_error(const c8b_blk* b) { return b ? b->err.c_str() : g_blkErr.c_str(); }

void c8b_blk_destroy(c8b_blk* b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->d_trig) cudaFree(b->d_trig);
    if (b->h_pin) cudaFreeHost(b->h_pin);
    for (Buf* q : { &b->d_a, &b->d_b, &b->d_c, &b->d_d }) if (q->p) cudaFree(q->p);
    if (b->ctx) c8b_destroy(b->ctx);
    delete b;
}

int c8b_blk_create(const c8b_cfg* cfg, int kind, c8b_blk** out)
{
    if (!out || kind < 0 || kind > C8B_BLK_DECODE) { g_blkErr = "c8b_blk_create: bad argument"; return C8B_ERR_ARG; }
    *out = nullptr;
    c8b_cfg c;
    memset(&c, 0, sizeof(c));
    if (cfg) c = *cfg;
    c.max_frames = 1;                                             // one frame in flight per block
    c.chunk_items = 1;
    c.decode_mode = 1;                                            // the latency decode kernel (one warp per frame)
    c8b_blk* b = new (std::nothrow) c8b_blk;
    if (!b) { g_blkErr = "out of memory"; return C8B_ERR_NOMEM; }
    b->kind = kind;
    b->device = c.device;
    int rc = c8b_create(&c, &b->ctx);
    if (rc) { g_blkErr = c8b_last_error(nullptr); delete b; return rc; }
    std::vector<uint8_t> blob(c8b_lut_size());
    if ((rc = c8b_lut_blob(blob.data(), blob.size())) || (rc = c8b_lut_load(b->ctx, blob.data(), blob.size()))) {
        g_blkErr = c8b_last_error(b->ctx);
        c8b_blk_destroy(b);
        return rc;
    }
    b->st = (cudaStream_t)c8b_stream(b->ctx);
    b->lut = c8b_ctx_lut(b->ctx);
    c8b::TrigState ts;
    c8b::trig_reset(ts);
    cudaError_t e = cudaMalloc(&b->d_trig, sizeof(ts));
    if (e == cudaSuccess) e = cudaMemcpy(b->d_trig, &ts, sizeof(ts), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        g_blkErr = std::string("c8b_blk_create: ") + cudaGetErrorString(e);
        c8b_blk_destroy(b);
        return C8B_ERR_CUDA;
    }
    *out = b;
    return C8B_OK;
}

int c8b_blk_work(c8b_blk* b, int noutput, const int* ninput, const void* const* in, void* const* out, const c8b_tag* in_tags,
                 int n_in_tags, int* consumed, int* produced, c8b_tag* out_tags, int out_tag_cap, int* n_out_tags, uint8_t* msg,
                 int msg_cap, int* msg_bytes)
{
    if (!b || !ninput || !in || !consumed || !produced || noutput < 0 || n_in_tags < 0 || (n_in_tags && !in_tags)) return C8B_ERR_ARG;
    int nin = 0, nout = 0;
    c8b_blk_ports(b->kind, &nin, &nout, nullptr, nullptr);
    for (int k = 0; k < nin; k++) if (ninput[k] < 0 || (ninput[k] && !in[k])) { b->err = "c8b_blk_work: bad input port"; return C8B_ERR_ARG; }
    for (int k = 0; k < nout; k++) if (noutput && (!out || !out[k])) { b->err = "c8b_blk_work: bad output port"; return C8B_ERR_ARG; }
    if (cudaSetDevice(b->device) != cudaSuccess) { b->err = "c8b_blk_work: cudaSetDevice failed"; return C8B_ERR_CUDA; }   // the calling
    c8b_blocks::WorkIO io;                                       // thread may have another device current (one block thread per GPU)
    io.noutput = noutput; io.ninput = ninput; io.in = in; io.out = out;
    io.in_tags = in_tags; io.n_in_tags = n_in_tags;
    io.out_tags = out_tags; io.out_tag_cap = out_tags ? out_tag_cap : 0;
    io.msg = msg; io.msg_cap = msg ? msg_cap : 0;
    int rc = C8B_OK;
    switch (b->kind) {
    case C8B_BLK_TRIGGER: rc = c8b_blocks::trigger_work(*b, io); break;
    case C8B_BLK_SYNC:    rc = c8b_blocks::sync_work(*b, b->sy, io); break;
    case C8B_BLK_SIGNAL:  rc = c8b_blocks::signal_work(*b, b->sg, 1, io); break;
    case C8B_BLK_SIGNAL2: rc = c8b_blocks::signal_work(*b, b->sg, 2, io); break;
    case C8B_BLK_DEMOD:   rc = c8b_blocks::demod_work(*b, b->dm, 1, io); break;
    case C8B_BLK_DEMOD2:  rc = c8b_blocks::demod_work(*b, b->dm, 2, io); break;
    default:              rc = c8b_blocks::decode_work(*b, b->dc, io); break;
    }
    if (rc == C8B_ERR_FULL) b->err = "c8b_blk_work: out_tags / msg too small";
    *consumed = io.consumed; *produced = io.produced;
    if (n_out_tags) *n_out_tags = io.n_out_tags;
    if (msg_bytes) *msg_bytes = io.msg_bytes;
    return rc;
}

}  // extern "C"
