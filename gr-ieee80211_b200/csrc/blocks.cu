// blocks.cu -- C ABI c8b_blk_*: the seven receive blocks of the reference, one scheduler call at a time
// (include/c80211b200.h).  The state machines are the host templates of blocks.h; this file is their sm_100a backend:
//   k_blk_trigger  : the trigger FSM (lib/trigger_impl.cc:59-117) continued from the state the previous call left in
//                    device memory; one warp, 32 samples per coalesced load, every lane steps the same FSM
//   k_blk_sync     : ltf_autoCorrelation + ltf_cfo of one trigger (lib/sync_impl.cc:155-196)
//   k_blk_signal   : L-SIG of one sync flag: 3 x (CFO rotation, DFT64), LS channel, 24-step Viterbi, parity / rate / length
//                    (lib/signal_impl.cc:108-162)
//   k_blk_cfo_copy : the S_COPY loop (lib/signal_impl.cc:164-192), one thread per sample
//   demod / demod2 / decode run the batch path's kernels on ONE frame (k_header_w / k_header2, k_demod / k_demod2,
//   k_viterbi) through c8b_demod / c8b_demod2 / c8b_decode.
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

__global__ void __launch_bounds__(32)
k_blk_trigger(c8b::TrigState* __restrict__ state, const float* __restrict__ in, int n, uint8_t* __restrict__ out)
{
    const int lane = threadIdx.x;
    c8b::TrigState s = *state;                                    // every lane carries the same FSM
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const float mine = i < n ? in[i] : 0.f;
        const int cnt = min(32, n - base);
        uint8_t o = 0;
        for (int k = 0; k < cnt; k++) {
            const float v = __shfl_sync(0xffffffffu, mine, k);
            const uint8_t r = c8b::trig_step(s, v);
            if (k == lane) o = r;
        }
        if (i < n) out[i] = o;
    }
    __syncwarp();
    if (lane == 0) *state = s;
}

__global__ void k_blk_sync(const float2* __restrict__ sig, float2 conj, SyncRes* __restrict__ res)
{
    if (threadIdx.x || blockIdx.x) return;
    const c8b::SyncOut o = c8b::sync_at(reinterpret_cast<const cf*>(sig), c8b::mk(conj.x, conj.y));
    res->ok = o.ok; res->mIndex = o.mIndex; res->rad = o.rad; res->snr = o.snr; res->rssi = o.rssi;
}

__global__ void k_blk_signal(const c8b_lut* __restrict__ lut, const float2* __restrict__ in, float rad, SignalRes* __restrict__ res)
{
    if (threadIdx.x || blockIdx.x) return;
    cf h[64];
    int mcs = 0, len = 0, nsamp = 0;
    const int ok = c8b::signal_at(lut, reinterpret_cast<const cf*>(in), rad, h, &mcs, &len, &nsamp);
    res->ok = ok; res->mcs = mcs; res->len = len; res->nsamp = nsamp;
    for (int k = 0; k < 64; k++) { res->chan[2 * k] = h[k].re; res->chan[2 * k + 1] = h[k].im; }
}

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
    // device scratch of the small kernels
    c8b::TrigState* d_trig = nullptr;
    SyncRes* d_sync = nullptr;
    SignalRes* d_sig = nullptr;
    Buf d_a, d_b, d_c, d_d;
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

    // ---- Ops backend of blocks.h ----
    int trigger(const float* in, int n, uint8_t* out)
    {
        int r;
        if ((r = grow(d_a, (size_t)n * sizeof(float))) || (r = grow(d_b, (size_t)n))) return r;
        BK(cudaMemcpyAsync(d_a.p, in, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
        k_blk_trigger<<<1, 32, 0, st>>>(d_trig, (const float*)d_a.p, n, (uint8_t*)d_b.p);
        BK(cudaGetLastError());
        BK(cudaMemcpyAsync(out, d_b.p, (size_t)n, cudaMemcpyDeviceToHost, st));
        BK(cudaStreamSynchronize(st));
        return C8B_OK;
    }
    int sync_at(const float* sig, const float conj[2], SyncRes* res)
    {
        int r;
        if ((r = grow(d_a, 240 * sizeof(float2)))) return r;
        BK(cudaMemcpyAsync(d_a.p, sig, 240 * sizeof(float2), cudaMemcpyHostToDevice, st));
        k_blk_sync<<<1, 32, 0, st>>>((const float2*)d_a.p, make_float2(conj[0], conj[1]), d_sync);
        BK(cudaGetLastError());
        BK(cudaMemcpyAsync(res, d_sync, sizeof(*res), cudaMemcpyDeviceToHost, st));
        BK(cudaStreamSynchronize(st));
        return C8B_OK;
    }
    int signal_at(const float* in, float rad, SignalRes* res)
    {
        int r;
        if ((r = grow(d_a, 224 * sizeof(float2)))) return r;
        BK(cudaMemcpyAsync(d_a.p, in, 224 * sizeof(float2), cudaMemcpyHostToDevice, st));
        k_blk_signal<<<1, 32, 0, st>>>(lut, (const float2*)d_a.p, rad, d_sig);
        BK(cudaGetLastError());
        BK(cudaMemcpyAsync(res, d_sig, sizeof(*res), cudaMemcpyDeviceToHost, st));
        BK(cudaStreamSynchronize(st));
        return C8B_OK;
    }
    int cfo_copy(const float* in0, const float* in1, float* out0, float* out1, int n, int copied, float rad)
    {
        const size_t bytes = (size_t)n * sizeof(float2);
        int r;
        if ((r = grow(d_a, bytes)) || (r = grow(d_b, bytes))) return r;
        if (in1 && ((r = grow(d_c, bytes)) || (r = grow(d_d, bytes)))) return r;
        BK(cudaMemcpyAsync(d_a.p, in0, bytes, cudaMemcpyHostToDevice, st));
        if (in1) BK(cudaMemcpyAsync(d_c.p, in1, bytes, cudaMemcpyHostToDevice, st));
        k_blk_cfo_copy<<<(n + 255) / 256, 256, 0, st>>>((const float2*)d_a.p, in1 ? (const float2*)d_c.p : nullptr, (float2*)d_b.p,
                                                        in1 ? (float2*)d_d.p : nullptr, n, copied, rad);
        BK(cudaGetLastError());
        BK(cudaMemcpyAsync(out0, d_b.p, bytes, cudaMemcpyDeviceToHost, st));
        if (in1) BK(cudaMemcpyAsync(out1, d_d.p, bytes, cudaMemcpyDeviceToHost, st));
        BK(cudaStreamSynchronize(st));
        return C8B_OK;
    }
    int demod(int nant, const float* iq0, const float* iq1, int n, c8b_frame* f, const float* chan, std::vector<float>* llr)
    {
        // soft bits of one frame: a short-GI symbol takes 72 samples (the rule of llr_stride_for in ctx.cu), 416 (one stream) /
        // 832 (two streams) soft bits per symbol
        const int64_t stride = std::max<int64_t>(((int64_t)f->nsamp / 72 + 1) * (nant == 2 ? 832 : 416), 1024);
        llr->assign((size_t)stride, 0.f);
        const int64_t off = 0;
        const int32_t len = n;
        const int rc = nant == 2 ? c8b_demod2(ctx, iq0, iq1, &off, &len, 1, f, chan, llr->data(), stride)
                                 : c8b_demod(ctx, iq0, &off, &len, 1, f, chan, llr->data(), stride);
        if (rc) err = c8b_last_error(ctx);
        return rc;
    }
    int decode(c8b_frame* f, const float* llr, int nllr, uint8_t* pdu, int pdu_cap)
    {
        const int rc = c8b_decode(ctx, llr, nllr, f, 1, pdu, pdu_cap, nullptr, 0);
        if (rc) err = c8b_last_error(ctx);
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
    if (b->d_sync) cudaFree(b->d_sync);
    if (b->d_sig) cudaFree(b->d_sig);
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
    if (e == cudaSuccess) e = cudaMalloc(&b->d_sync, sizeof(SyncRes));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_sig, sizeof(SignalRes));
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
