// ctx.cu -- context, scratch management and the C ABI of libc80211b200.so (include/c80211b200.h).
// Host code only drives: every sample, soft bit and decoded byte is produced by the sm_100a kernels
// in k_*.cu.  There is no CPU path: without a CUDA device c8b_create fails.
#include <new>
#include <string>
#include <vector>
#include <string.h>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached

#include "common.cuh"

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct c8b_ctx {
    c8b_cfg cfg;
    int device = 0, numSM = 0;
    cudaStream_t st = nullptr, stCopy = nullptr, stVit = nullptr;
    cudaEvent_t evFront[2] = { nullptr, nullptr }, evVit[2] = { nullptr, nullptr };
    unsigned chunkSeq = 0;
    bool overlap = true;     // Viterbi of chunk k on its own stream, concurrent with the front end of chunk k+1
    std::string err;
    c8b_lut* d_lut = nullptr;
    bool lutLoaded = false;
    unsigned* d_counter = nullptr;
    // scratch (grown on demand)
    DevBuf iq, iqf, iq1, mask, llrB, tp, preac, preconj, trig, off, len, frames, chan, hinv, w2, llr, surv, pdu, scram;
    int survWarps = 0;
    // live-stream session (c8b_stream_*): a device-resident window of the capture, ping-pong compacted
    struct Stream {
        bool open = false;
        int nant = 1, cur = 0;
        int64_t cap = 0;        // window capacity, samples
        int64_t base = 0;       // absolute stream index of window sample 0
        int64_t fill = 0;       // samples in the window
        int32_t from = 0;       // first window sample the trigger FSM sees
        int64_t posAbs = 0;     // signal block's consumed-until, absolute
        int64_t overruns = 0;   // windows dropped because nothing in them could be decided
    } strm;
    DevBuf sw[2][2], scan;      // [antenna][ping-pong]
    // one frame per call (the per-block entry points, blocks.cu): packed input / output blocks, pinned + device
    struct OneSlot { void* host = nullptr; size_t hostCap = 0; DevBuf dev; cudaEvent_t ev = nullptr; uint32_t seq = 0; int64_t aux = 0, aux2 = 0; };
    uint32_t* oneFlagH = nullptr;   // C8B_ONE_SLOTS completion words in mapped pinned memory (c8b_launch_flag / c8b_wait_flag)
    uint32_t* oneFlagD = nullptr;
    uint32_t oneSeq = 0;
    OneSlot one[C8B_ONE_SLOTS];     // staging slots of the asynchronous one-frame ops (demod: 0 and 1; decode: all, round robin)
    // pinned read-back staging of a stream pass: frame records + scan result, then the PDU area of the frames taken
    void* hStage = nullptr;
    size_t hStageCap = 0;
    DevBuf cand;                // candidate records of the multi-frame detect path
    DevBuf txf, txplan, txpsdu, txiq;   // transmit synthesiser: descriptors, plans, staged PSDU bytes / samples
    c8b_scan* scanDev = nullptr;   // non-null while run_chunk serves a stream window
    // item table (off / len) resident on the device: a pinned host mirror of what ctx->off / ctx->len hold for items
    // [tabLo, tabHi), so that a batch call with an unchanged table (a bench step, a monitor re-scanning the same arena)
    // uploads nothing, and a changed slice goes up asynchronously right before the chunk that needs it
    int64_t* tabOff = nullptr;
    int32_t* tabLen = nullptr;
    size_t tabCap = 0;
    int tabLo = 0, tabHi = 0;
    cudaEvent_t evTab = nullptr;
    // timing
    bool timing = false;
    double ms[C8B_K_COUNT] = { 0 };
    int64_t launches[C8B_K_COUNT] = { 0 };
    struct Pending { int k; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> evPool;
};

static std::string g_createErr;

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
            return C8B_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

static int ensure(c8b_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return C8B_OK;
    if (b.p) { cudaDeviceSynchronize(); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e);
        return C8B_ERR_NOMEM;
    }
    b.cap = want;
    return C8B_OK;
}

#define EN(buf, bytes)                                   \
    do {                                                 \
        int r_ = ensure(ctx, ctx->buf, (size_t)(bytes)); \
        if (r_) return r_;                               \
    } while (0)

// ---- stage timing with CUDA events on the launching stream ---------------------------------------
static cudaEvent_t ev_get(c8b_ctx* ctx)
{
    if (!ctx->evPool.empty()) { cudaEvent_t e = ctx->evPool.back(); ctx->evPool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
// One stage of the chain (tracing): an NVTX range around the launches -- the five ranges presiso / detect /
// header / demod / viterbi are what a timeline shows per chunk -- and, when timing is on, a CUDA event pair on the stream.
static const char* const kStageName[C8B_K_COUNT] = { "c8b presiso", "c8b detect (trigger+sync+signal)", "c8b header", "c8b demod", "c8b viterbi" };
struct StageTimer {
    c8b_ctx* ctx; int k; cudaStream_t s; cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(c8b_ctx* c, int kk, cudaStream_t ss = nullptr) : ctx(c), k(kk), s(ss ? ss : c->st)
    {
        nvtxRangePushA(kStageName[k]);
        if (ctx->timing) { a = ev_get(ctx); b = ev_get(ctx); cudaEventRecord(a, s); }
    }
    ~StageTimer()
    {
        if (ctx->timing) { cudaEventRecord(b, s); ctx->pending.push_back({ k, a, b }); }
        nvtxRangePop();
    }
};
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
static void timing_collect(c8b_ctx* ctx)
{
    for (auto& p : ctx->pending) {
        float ms = 0.f;
        cudaEventSynchronize(p.b);
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { ctx->ms[p.k] += ms; ctx->launches[p.k]++; }
        ctx->evPool.push_back(p.a);
        ctx->evPool.push_back(p.b);
    }
    ctx->pending.clear();
}

extern "C" {

int c8b_abi_version(void) { return C8B_ABI_VERSION; }

int c8b_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* c8b_last_error(const c8b_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createErr.c_str(); }

int c8b_create(const c8b_cfg* cfg, c8b_ctx** out)
{
    if (!out) return C8B_ERR_ARG;
    *out = nullptr;
    int n = c8b_device_count();
    if (n <= 0) { g_createErr = "no CUDA device visible: libc80211b200 has no CPU path"; return C8B_ERR_NO_DEVICE; }
    c8b_ctx* ctx = new (std::nothrow) c8b_ctx();
    if (!ctx) return C8B_ERR_NOMEM;
    memset(&ctx->cfg, 0, sizeof(ctx->cfg));
    if (cfg) ctx->cfg = *cfg;
    if (ctx->cfg.max_frames <= 0) ctx->cfg.max_frames = 1;
    ctx->device = ctx->cfg.device;
    if (ctx->device < 0 || ctx->device >= n) { g_createErr = "bad device ordinal"; delete ctx; return C8B_ERR_ARG; }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stCopy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stVit, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaEventCreateWithFlags(&ctx->evFront[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evVit[k], cudaEventDisableTiming);
    }
    ctx->overlap = ctx->cfg.no_overlap == 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->numSM, cudaDevAttrMultiProcessorCount, ctx->device);
    if (ctx->cfg.chunk_items <= 0) ctx->cfg.chunk_items = c8b_viterbi_tp_wave(e == cudaSuccess && ctx->numSM > 0 ? ctx->numSM : 148);
    if (e == cudaSuccess) e = c8b_viterbi_prepare();          // function attributes are per device: set on THIS context's device
    if (e == cudaSuccess) e = c8b_viterbi_tp_prepare();
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_lut, sizeof(c8b_lut));
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_counter, 64);
    if (e != cudaSuccess) { g_createErr = std::string("c8b_create: ") + cudaGetErrorString(e); delete ctx; return C8B_ERR_CUDA; }
    *out = ctx;
    return C8B_OK;
}

void c8b_destroy(c8b_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    timing_collect(ctx);
    for (auto e : ctx->evPool) cudaEventDestroy(e);
    for (int k = 0; k < 2; k++) { if (ctx->evFront[k]) cudaEventDestroy(ctx->evFront[k]); if (ctx->evVit[k]) cudaEventDestroy(ctx->evVit[k]); }
    if (ctx->stVit) cudaStreamDestroy(ctx->stVit);
    DevBuf* bufs[] = { &ctx->iq, &ctx->iqf, &ctx->iq1, &ctx->w2, &ctx->mask, &ctx->llrB, &ctx->tp, &ctx->preac, &ctx->preconj, &ctx->trig, &ctx->off, &ctx->len, &ctx->frames, &ctx->chan,
                       &ctx->hinv, &ctx->llr, &ctx->surv, &ctx->pdu, &ctx->scram, &ctx->scan, &ctx->sw[0][0], &ctx->sw[0][1],
                       &ctx->sw[1][0], &ctx->sw[1][1], &ctx->txf, &ctx->txplan, &ctx->txpsdu, &ctx->txiq, &ctx->cand };
    for (auto b : bufs) if (b->p) cudaFree(b->p);
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->hStage) cudaFreeHost(ctx->hStage);
    if (ctx->oneFlagH) cudaFreeHost(ctx->oneFlagH);
    for (auto& sl : ctx->one) { if (sl.host) cudaFreeHost(sl.host); if (sl.dev.p) cudaFree(sl.dev.p); if (sl.ev) cudaEventDestroy(sl.ev); }
    if (ctx->tabOff) cudaFreeHost(ctx->tabOff);
    if (ctx->tabLen) cudaFreeHost(ctx->tabLen);
    if (ctx->evTab) cudaEventDestroy(ctx->evTab);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    if (ctx->stCopy) cudaStreamDestroy(ctx->stCopy);
    delete ctx;
}

void* c8b_stream(c8b_ctx* ctx) { return ctx ? (void*)ctx->st : nullptr; }

int c8b_sync(c8b_ctx* ctx)
{
    if (!ctx) return C8B_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaStreamSynchronize(ctx->stVit));
    return C8B_OK;
}

// ---- LUT ------------------------------------------------------------------------------------------
size_t c8b_lut_size(void) { return sizeof(c8b_lut); }

int c8b_lut_blob(void* buf, size_t cap)
{
    if (!buf || cap < sizeof(c8b_lut)) return C8B_ERR_ARG;
    c8b_lut_build(reinterpret_cast<c8b_lut*>(buf));
    return C8B_OK;
}

static int lut_check(c8b_ctx* ctx, const c8b_lut* h, size_t n)
{
    if (n != sizeof(c8b_lut) || h->magic != C8B_LUT_MAGIC || h->version != C8B_LUT_VERSION || h->bytes != sizeof(c8b_lut)) {
        ctx->err = "LUT blob: wrong size/magic/version";
        return C8B_ERR_LUT;
    }
    return C8B_OK;
}

int c8b_lut_load(c8b_ctx* ctx, const void* blob, size_t n)
{
    if (!ctx || !blob) return C8B_ERR_ARG;
    int r = lut_check(ctx, reinterpret_cast<const c8b_lut*>(blob), n);
    if (r) return r;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->d_lut, blob, n, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->lutLoaded = true;
    return C8B_OK;
}

int c8b_lut_load_dev(c8b_ctx* ctx, const void* d_blob, size_t n)
{
    if (!ctx || !d_blob || n != sizeof(c8b_lut)) return C8B_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    c8b_lut hdr;
    CK(cudaMemcpy(&hdr, d_blob, sizeof(hdr), cudaMemcpyDeviceToHost));
    int r = lut_check(ctx, &hdr, n);
    if (r) return r;
    CK(cudaMemcpyAsync(ctx->d_lut, d_blob, n, cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->lutLoaded = true;
    return C8B_OK;
}

const c8b_lut* c8b_ctx_lut(const c8b_ctx* ctx) { return ctx && ctx->lutLoaded ? ctx->d_lut : nullptr; }

static int need_lut(c8b_ctx* ctx)
{
    if (!ctx->lutLoaded) { ctx->err = "LUT not loaded: call c8b_lut_load first"; return C8B_ERR_LUT; }
    return C8B_OK;
}

// ---- timing ---------------------------------------------------------------------------------------
int c8b_timing_enable(c8b_ctx* ctx, int on)
{
    if (!ctx) return C8B_ERR_ARG;
    ctx->timing = on != 0;
    return C8B_OK;
}

int c8b_timing_read(c8b_ctx* ctx, double ms[C8B_K_COUNT], int64_t launches[C8B_K_COUNT], int reset)
{
    if (!ctx) return C8B_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    timing_collect(ctx);
    for (int i = 0; i < C8B_K_COUNT; i++) {
        if (ms) ms[i] = ctx->ms[i];
        if (launches) launches[i] = ctx->launches[i];
        if (reset) { ctx->ms[i] = 0; ctx->launches[i] = 0; }
    }
    return C8B_OK;
}

// ---- decode stage ---------------------------------------------------------------------------------

static int ensure_surv(c8b_ctx* ctx)
{
    int grid = c8b_viterbi_max_grid(ctx->numSM);
    int warps = grid * C8B_VIT_WARPS;
    if (ctx->survWarps >= warps) return C8B_OK;
    EN(surv, (size_t)warps * C8B_VIT_WARP_SLOTS * sizeof(uint2));   // two frames per warp + decoded group words
    ctx->survWarps = warps;
    return C8B_OK;
}

// Which decode kernel: one thread per frame (throughput, large batches) or one warp per frame pair (latency).
static bool use_thread_per_frame(const c8b_ctx* ctx, int nframes)
{
    if (ctx->cfg.decode_mode == 1) return false;
    if (ctx->cfg.decode_mode == 2) return true;
    return nframes >= 4096;
}

static int launch_decode(c8b_ctx* ctx, c8b_frame* d_frames, int n, const float* d_llr, int64_t nllr, uint8_t* d_pdu, int64_t pdu_stride,
                         uint8_t* d_scram, int64_t scram_stride, int grid, cudaStream_t st)
{
    if (use_thread_per_frame(ctx, n)) {
        EN(tp, c8b_viterbi_tp_scratch_bytes(ctx->numSM, n));
        c8b_launch_viterbi_tp(ctx->d_lut, d_frames, n, d_llr, nllr, ctx->tp.p, ctx->numSM, d_pdu, pdu_stride, d_scram, scram_stride, st);
    } else {
        int r = ensure_surv(ctx);
        if (r) return r;
        c8b_launch_viterbi(ctx->d_lut, d_frames, n, d_llr, nllr, (uint2*)ctx->surv.p, ctx->survWarps, d_pdu, pdu_stride, d_scram, scram_stride,
                           ctx->d_counter, grid, st);
    }
    return C8B_OK;
}

// device-resident decode of d_frames[0..n): LLR arena d_llr (nllr floats), PDUs to d_pdu
static int decode_dev(c8b_ctx* ctx, c8b_frame* d_frames, int n, const float* d_llr, int64_t nllr, uint8_t* d_pdu,
                      int64_t pdu_stride, uint8_t* d_scram, int64_t scram_stride)
{
    {
        StageTimer tm(ctx, C8B_K_VITERBI);
        int r = launch_decode(ctx, d_frames, n, d_llr, nllr, d_pdu, pdu_stride, d_scram, scram_stride, c8b_viterbi_max_grid(ctx->numSM), ctx->st);
        if (r) return r;
    }
    c8b_launch_ndp(d_frames, n, d_llr, nllr, d_pdu, pdu_stride, ctx->st);    // VHT NDP channel reports (lib/decode_impl.cc:100-121)
    CK(cudaGetLastError());
    return C8B_OK;
}

int c8b_decode(c8b_ctx* ctx, const float* h_llr, int64_t nllr, c8b_frame* frames, int nframes, uint8_t* h_pdu,
               int64_t pdu_stride, uint8_t* h_scram, int64_t scram_stride)
{
    if (!ctx || !h_llr || !frames || nframes < 0 || nllr < 0 || !h_pdu || pdu_stride <= 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    if (nframes == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    EN(llr, (size_t)(nllr + 4) * sizeof(float));
    EN(frames, (size_t)nframes * sizeof(c8b_frame));
    EN(pdu, (size_t)nframes * pdu_stride);
    if (h_scram) EN(scram, (size_t)nframes * scram_stride);
    CK(cudaMemcpyAsync(ctx->llr.p, h_llr, (size_t)nllr * sizeof(float), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->frames.p, frames, (size_t)nframes * sizeof(c8b_frame), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(ctx->pdu.p, 0, (size_t)nframes * pdu_stride, ctx->st));
    if (h_scram) CK(cudaMemsetAsync(ctx->scram.p, 0, (size_t)nframes * scram_stride, ctx->st));
    r = decode_dev(ctx, (c8b_frame*)ctx->frames.p, nframes, (const float*)ctx->llr.p, nllr, (uint8_t*)ctx->pdu.p, pdu_stride,
                   h_scram ? (uint8_t*)ctx->scram.p : nullptr, scram_stride);
    if (r) return r;
    CK(cudaMemcpyAsync(frames, ctx->frames.p, (size_t)nframes * sizeof(c8b_frame), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(h_pdu, ctx->pdu.p, (size_t)nframes * pdu_stride, cudaMemcpyDeviceToHost, ctx->st));
    if (h_scram) CK(cudaMemcpyAsync(h_scram, ctx->scram.p, (size_t)nframes * scram_stride, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

}  // extern "C"

// ---- front end + demod + whole chain -----------------------------------------------------------------
struct ChunkPlan { int64_t base = 0, end = 0; int maxLen = 0; };

static ChunkPlan plan(const int64_t* off, const int32_t* len, int n)
{
    ChunkPlan p;
    if (n <= 0) return p;
    p.base = off[0]; p.end = off[0] + len[0];
    for (int i = 0; i < n; i++) {
        if (off[i] < p.base) p.base = off[i];
        if (off[i] + len[i] > p.end) p.end = off[i] + len[i];
        if (len[i] > p.maxLen) p.maxLen = len[i];
    }
    return p;
}

static int64_t llr_stride_for(int maxLen)
{
    // every DATA symbol takes >= 72 samples and starts >= 224 samples after the sync index
    int64_t nsym = maxLen / 72 + 1;
    if (nsym > 1366) nsym = 1366;                       // lib/decode_impl.h:36: trellis <= 32782 -> <= 1366 BPSK symbols
    return nsym * 416;
}

static int check_items(c8b_ctx* ctx, const int64_t* off, const int32_t* len, int n)
{
    for (int i = 0; i < n; i++)
        if (off[i] < 0 || len[i] < 0) { ctx->err = "negative item offset/length"; return C8B_ERR_ARG; }
    return C8B_OK;
}

// items [b, e) of a device-resident capture: all stages, results into d_frames[b..e) / d_pdu
static int run_chunk(c8b_ctx* ctx, const float2* d_iq, const int64_t* d_off, const int32_t* d_len, const int64_t* h_off,
                     const int32_t* h_len, int b, int e, c8b_frame* d_frames, uint8_t* d_pdu, int64_t pdu_stride, int64_t iqShift,
                     const float2* d_iq1 = nullptr)
{
    NvtxRange chunkRange("c8b chunk");
    const int n = e - b, maxf = ctx->cfg.max_frames;
    const size_t ns = (size_t)n * maxf;                          // frame slots of this chunk
    const ChunkPlan pl = plan(h_off + b, h_len + b, n);
    const int64_t span = pl.end - pl.base;
    const int64_t llrStride = llr_stride_for(pl.maxLen) * (d_iq1 ? 2 : 1);
    if (d_iq1) EN(w2, ns * 264 * sizeof(float2));
    const int maskStride = (pl.maxLen + 31) / 32 + 1;
    EN(preac, (size_t)(span + 64) * sizeof(float));
    EN(mask, (size_t)n * maskStride * sizeof(uint32_t));
    EN(chan, ns * 64 * sizeof(float2));
    EN(hinv, ns * 64 * sizeof(float2));
    const int par = (int)(ctx->chunkSeq++ & 1u);
    const bool ov = ctx->overlap;
    DevBuf& llrBuf = (ov && par) ? ctx->llrB : ctx->llr;       // double-buffered when the Viterbi runs on its own stream
    {
        int r_ = ensure(ctx, llrBuf, ns * llrStride * sizeof(float));
        if (r_) return r_;
    }
    int r = C8B_OK;
    if (ov) cudaStreamWaitEvent(ctx->st, ctx->evVit[par], 0);   // the Viterbi pass that last read this LLR buffer
    // iqShift: the device buffer holds the capture from sample iqShift on (host-staged chunks)
    const float2* iq = d_iq - iqShift;
    const float2* iq1 = d_iq1 ? d_iq1 - iqShift : nullptr;
    {
        StageTimer tm(ctx, C8B_K_PRESISO);
        c8b_launch_presiso(iq, d_off + b, d_len + b, n, pl.maxLen, pl.base, (float*)ctx->preac.p, nullptr, (uint32_t*)ctx->mask.p, maskStride,
                           ctx->st);
    }
    {
        // few long items with many frames each (a capture, a stream window): the per-trigger work runs in parallel
        const bool multi = ctx->cfg.frontend_mode == 0 && maxf > 1 && n <= 64;
        if (multi) EN(cand, c8b_detect_multi_scratch(n, pl.maxLen));
        StageTimer tm(ctx, C8B_K_DETECT);
        if (multi)
            c8b_launch_detect_multi(ctx->d_lut, iq, d_off + b, d_len + b, n, b, maxf, pl.base, (const float*)ctx->preac.p,
                                    (const uint32_t*)ctx->mask.p, maskStride, d_frames + (size_t)b * maxf, (float2*)ctx->chan.p, ctx->scanDev,
                                    ctx->cand.p, pl.maxLen, ctx->st);
        else
            (ctx->cfg.frontend_mode == 1 ? c8b_launch_detect : c8b_launch_detect_w)(
                ctx->d_lut, iq, d_off + b, d_len + b, n, b, maxf, pl.base, (const float*)ctx->preac.p, (const uint32_t*)ctx->mask.p, maskStride,
                d_frames + (size_t)b * maxf, (float2*)ctx->chan.p, ctx->scanDev, ctx->st);
    }
    {
        StageTimer tm(ctx, C8B_K_HEADER);
        if (iq1)
            (ctx->cfg.frontend_mode == 1 ? c8b_launch_header2 : c8b_launch_header2_w)(
                ctx->d_lut, iq, iq1, d_off + b, n, maxf, ctx->cfg.mmse, d_frames + (size_t)b * maxf, (const float2*)ctx->chan.p,
                (float2*)ctx->hinv.p, (float2*)ctx->w2.p, llrStride, ctx->st);
        else
            (ctx->cfg.frontend_mode == 1 ? c8b_launch_header : c8b_launch_header_w)(
                ctx->d_lut, iq, d_off + b, n, maxf, ctx->cfg.mupos, d_frames + (size_t)b * maxf, (const float2*)ctx->chan.p,
                (float2*)ctx->hinv.p, llrStride, (float*)llrBuf.p, ctx->st);
    }
    {
        StageTimer tm(ctx, C8B_K_DEMOD);
        const int maxSym = (int)(llr_stride_for(pl.maxLen) / 416);
        c8b_launch_demod(ctx->d_lut, iq, d_off + b, n, maxf, maxSym, d_frames + (size_t)b * maxf, (const float2*)ctx->hinv.p, (float*)llrBuf.p,
                         ctx->st);
        if (iq1)
            c8b_launch_demod2(ctx->d_lut, iq, iq1, d_off + b, n, maxf, maxSym, d_frames + (size_t)b * maxf, (const float2*)ctx->w2.p,
                              (float*)llrBuf.p, ctx->st);
    }
    cudaStream_t sv = ctx->st;
    if (ov) {
        sv = ctx->stVit;
        cudaEventRecord(ctx->evFront[par], ctx->st);
        cudaStreamWaitEvent(sv, ctx->evFront[par], 0);
    }
    {
        StageTimer tm(ctx, C8B_K_VITERBI, sv);
        // with the front end of the next chunk co-resident, one CTA per SM less (registers / shared memory)
        const int grid = ov ? ctx->numSM * 4 : c8b_viterbi_max_grid(ctx->numSM);
        r = launch_decode(ctx, d_frames + (size_t)b * maxf, (int)ns, (const float*)llrBuf.p, (int64_t)ns * llrStride,
                          d_pdu + (size_t)b * maxf * pdu_stride, pdu_stride, nullptr, 0, grid, sv);
        if (r) return r;
    }
    c8b_launch_ndp(d_frames + (size_t)b * maxf, (int)ns, (const float*)llrBuf.p, (int64_t)ns * llrStride, d_pdu + (size_t)b * maxf * pdu_stride,
                   pdu_stride, sv);                                  // VHT NDP channel reports (lib/decode_impl.cc:100-121)
    if (ov) cudaEventRecord(ctx->evVit[par], sv);
    CK(cudaGetLastError());
    return C8B_OK;
}

// make the ctx stream wait for every Viterbi pass in flight (end of a batched call)
static void join_viterbi(c8b_ctx* ctx)
{
    if (!ctx->overlap) return;
    cudaStreamWaitEvent(ctx->st, ctx->evVit[0], 0);
    cudaStreamWaitEvent(ctx->st, ctx->evVit[1], 0);
}

// items [b, e) of the caller's table -> ctx->off / ctx->len at the same indices, unless the mirror says they are there already
static int upload_items_range(c8b_ctx* ctx, const int64_t* off, const int32_t* len, int n, int b, int e)
{
    if ((size_t)n > ctx->tabCap || ctx->off.cap < (size_t)n * sizeof(int64_t) || ctx->len.cap < (size_t)n * sizeof(int32_t)) {
        if (ctx->evTab) cudaEventSynchronize(ctx->evTab);
        EN(off, (size_t)n * sizeof(int64_t));
        EN(len, (size_t)n * sizeof(int32_t));
        if (ctx->hStage) cudaFreeHost(ctx->hStage);
    if (ctx->oneFlagH) cudaFreeHost(ctx->oneFlagH);
    for (auto& sl : ctx->one) { if (sl.host) cudaFreeHost(sl.host); if (sl.dev.p) cudaFree(sl.dev.p); if (sl.ev) cudaEventDestroy(sl.ev); }
    if (ctx->tabOff) cudaFreeHost(ctx->tabOff);
        if (ctx->tabLen) cudaFreeHost(ctx->tabLen);
        ctx->tabOff = nullptr; ctx->tabLen = nullptr; ctx->tabCap = 0; ctx->tabLo = ctx->tabHi = 0;
        const size_t cap = (size_t)n + (size_t)n / 8 + 64;
        CK(cudaHostAlloc((void**)&ctx->tabOff, cap * sizeof(int64_t), cudaHostAllocDefault));
        CK(cudaHostAlloc((void**)&ctx->tabLen, cap * sizeof(int32_t), cudaHostAllocDefault));
        ctx->tabCap = cap;
        if (!ctx->evTab) CK(cudaEventCreateWithFlags(&ctx->evTab, cudaEventDisableTiming));
    }
    const size_t m = (size_t)(e - b);
    if (ctx->tabLo <= b && e <= ctx->tabHi && memcmp(ctx->tabOff + b, off + b, m * sizeof(int64_t)) == 0 &&
        memcmp(ctx->tabLen + b, len + b, m * sizeof(int32_t)) == 0)
        return C8B_OK;
    cudaEventSynchronize(ctx->evTab);                              // an earlier upload may still be reading the mirror
    memcpy(ctx->tabOff + b, off + b, m * sizeof(int64_t));
    memcpy(ctx->tabLen + b, len + b, m * sizeof(int32_t));
    CK(cudaMemcpyAsync((int64_t*)ctx->off.p + b, ctx->tabOff + b, m * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync((int32_t*)ctx->len.p + b, ctx->tabLen + b, m * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaEventRecord(ctx->evTab, ctx->st));
    if (b <= ctx->tabHi && e >= ctx->tabLo && ctx->tabHi > ctx->tabLo) { if (b < ctx->tabLo) ctx->tabLo = b; if (e > ctx->tabHi) ctx->tabHi = e; }
    else { ctx->tabLo = b; ctx->tabHi = e; }
    return C8B_OK;
}

static int upload_items(c8b_ctx* ctx, const int64_t* off, const int32_t* len, int n)
{
    ctx->tabLo = ctx->tabHi = 0;                                   // the staged entry points overwrite the table: the mirror is void
    EN(off, (size_t)n * sizeof(int64_t));
    EN(len, (size_t)n * sizeof(int32_t));
    CK(cudaMemcpyAsync(ctx->off.p, off, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->len.p, len, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->st));
    return C8B_OK;
}

extern "C" {

static int rx_batch_dev_async2(c8b_ctx* ctx, const float* d_iq, const float* d_iq1, const int64_t* off, const int32_t* len, int nitems,
                               c8b_frame* d_frames, uint8_t* d_pdu, int64_t pdu_stride)
{
    if (!ctx || !d_iq || !off || !len || nitems < 0 || !d_frames || !d_pdu || pdu_stride <= 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    if (nitems == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    if ((r = check_items(ctx, off, len, nitems))) return r;
    CK(cudaMemsetAsync(d_frames, 0, (size_t)nitems * ctx->cfg.max_frames * sizeof(c8b_frame), ctx->st));
    const int cs = ctx->cfg.chunk_items;
    for (int b = 0; b < nitems; b += cs) {
        const int e = b + cs < nitems ? b + cs : nitems;
        if ((r = upload_items_range(ctx, off, len, nitems, b, e))) return r;
        r = run_chunk(ctx, (const float2*)d_iq, (const int64_t*)ctx->off.p, (const int32_t*)ctx->len.p, off, len, b, e, d_frames, d_pdu,
                      pdu_stride, 0, (const float2*)d_iq1);
        if (r) return r;
    }
    join_viterbi(ctx);
    return C8B_OK;
}

int c8b_rx_batch_dev_async(c8b_ctx* ctx, const float* d_iq, const int64_t* off, const int32_t* len, int nitems, c8b_frame* d_frames,
                           uint8_t* d_pdu, int64_t pdu_stride)
{
    return rx_batch_dev_async2(ctx, d_iq, nullptr, off, len, nitems, d_frames, d_pdu, pdu_stride);
}

int c8b_rx_batch2_dev(c8b_ctx* ctx, const float* d_iq0, const float* d_iq1, const int64_t* off, const int32_t* len, int nitems,
                      c8b_frame* frames, uint8_t* pdu, int64_t pdu_stride)
{
    if (!ctx || !d_iq1 || !frames || !pdu || nitems < 0 || pdu_stride <= 0) return C8B_ERR_ARG;
    if (nitems == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t nsl = (size_t)nitems * (ctx->cfg.max_frames > 0 ? ctx->cfg.max_frames : 1);
    EN(frames, nsl * sizeof(c8b_frame));
    EN(pdu, nsl * pdu_stride);
    int r = rx_batch_dev_async2(ctx, d_iq0, d_iq1, off, len, nitems, (c8b_frame*)ctx->frames.p, (uint8_t*)ctx->pdu.p, pdu_stride);
    if (r) return r;
    CK(cudaMemcpyAsync(frames, ctx->frames.p, nsl * sizeof(c8b_frame), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(pdu, ctx->pdu.p, nsl * pdu_stride, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

int c8b_rx_batch_dev(c8b_ctx* ctx, const float* d_iq, const int64_t* off, const int32_t* len, int nitems, c8b_frame* frames,
                     uint8_t* pdu, int64_t pdu_stride)
{
    if (!ctx || !frames || !pdu || nitems < 0 || pdu_stride <= 0) return C8B_ERR_ARG;
    if (nitems == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t nsl = (size_t)nitems * (ctx->cfg.max_frames > 0 ? ctx->cfg.max_frames : 1);
    EN(frames, nsl * sizeof(c8b_frame));
    EN(pdu, nsl * pdu_stride);
    int r = c8b_rx_batch_dev_async(ctx, d_iq, off, len, nitems, (c8b_frame*)ctx->frames.p, (uint8_t*)ctx->pdu.p, pdu_stride);
    if (r) return r;
    CK(cudaMemcpyAsync(frames, ctx->frames.p, nsl * sizeof(c8b_frame), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(pdu, ctx->pdu.p, nsl * pdu_stride, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

// Host IQ: chunks are staged into two device buffers on the copy stream while the previous chunk is
// processed on the compute stream; results of each chunk go back as soon as its Viterbi kernel ends.
// sc16 = 1: h_iq holds interleaved int16 (I, Q) pairs -- the wire format of a USRP / UHD stream (sc16), half the bytes of
// fc32 over PCIe; each chunk is widened on the device to exactly the floats UHD's own converter hands the reference
// (x * (1 / 32768)) into one fc32 buffer the front end reads (the widening of chunk k+1 follows the front end of chunk k
// on the compute stream, so one buffer is enough; the raw chunks are double-buffered like the fc32 ones).
static int rx_batch_host(c8b_ctx* ctx, const void* h_iq, const float* h_iq1, int sc16, const int64_t* off, const int32_t* len, int nitems,
                         c8b_frame* frames, uint8_t* pdu, int64_t pdu_stride)
{
    if (!ctx || !h_iq || !off || !len || nitems < 0 || !frames || !pdu || pdu_stride <= 0 || (sc16 && h_iq1)) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    if (nitems == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    if ((r = check_items(ctx, off, len, nitems))) return r;
    const size_t maxf = (size_t)ctx->cfg.max_frames;
    EN(frames, (size_t)nitems * maxf * sizeof(c8b_frame));
    EN(pdu, (size_t)nitems * maxf * pdu_stride);
    CK(cudaMemsetAsync(ctx->frames.p, 0, (size_t)nitems * maxf * sizeof(c8b_frame), ctx->st));
    const int cs = ctx->cfg.chunk_items;
    const int nchunks = (nitems + cs - 1) / cs;
    // size the two staging buffers for the largest chunk span
    int64_t maxSpan = 0;
    for (int c = 0; c < nchunks; c++) {
        const int b = c * cs, e = b + cs < nitems ? b + cs : nitems;
        const ChunkPlan pl = plan(off + b, len + b, e - b);
        if (pl.end - pl.base > maxSpan) maxSpan = pl.end - pl.base;
    }
    const size_t sampBytes = sc16 ? sizeof(short2) : sizeof(float2);
    const int64_t slot = (maxSpan + 16 + 3) & ~(int64_t)3;        // samples per staging buffer (16-byte multiples for sc16 too)
    EN(iq, (size_t)2 * slot * sampBytes);
    if (sc16) EN(iqf, (size_t)slot * sizeof(float2));
    if (h_iq1) EN(iq1, (size_t)2 * slot * sizeof(float2));
    char* buf[2] = { (char*)ctx->iq.p, (char*)ctx->iq.p + (size_t)slot * sampBytes };
    float2* buf1[2] = { h_iq1 ? (float2*)ctx->iq1.p : nullptr, h_iq1 ? (float2*)ctx->iq1.p + slot : nullptr };
    cudaEvent_t copied[2], freed[2];
    for (int k = 0; k < 2; k++) { cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&freed[k], cudaEventDisableTiming); }
    int rc = C8B_OK;
    auto stage = [&](int c) -> cudaError_t {
        const int b = c * cs, e = b + cs < nitems ? b + cs : nitems, k = c & 1;
        const ChunkPlan pl = plan(off + b, len + b, e - b);
        if (c >= 2) cudaStreamWaitEvent(ctx->stCopy, freed[k], 0);
        cudaError_t er = cudaMemcpyAsync(buf[k], reinterpret_cast<const char*>(h_iq) + (size_t)pl.base * sampBytes, (size_t)(pl.end - pl.base) * sampBytes,
                                         cudaMemcpyHostToDevice, ctx->stCopy);
        if (h_iq1 && er == cudaSuccess)
            er = cudaMemcpyAsync(buf1[k], reinterpret_cast<const float2*>(h_iq1) + pl.base, (size_t)(pl.end - pl.base) * sizeof(float2),
                                 cudaMemcpyHostToDevice, ctx->stCopy);
        cudaEventRecord(copied[k], ctx->stCopy);
        return er;
    };
    cudaError_t er = stage(0);
    for (int c = 0; c < nchunks && er == cudaSuccess && rc == C8B_OK; c++) {
        const int b = c * cs, e = b + cs < nitems ? b + cs : nitems, k = c & 1;
        if (c + 1 < nchunks) er = stage(c + 1);
        const ChunkPlan pl = plan(off + b, len + b, e - b);
        cudaStreamWaitEvent(ctx->st, copied[k], 0);
        const float2* src = reinterpret_cast<const float2*>(buf[k]);
        if (sc16) {
            c8b_launch_sc16_to_fc32(reinterpret_cast<const short2*>(buf[k]), (float2*)ctx->iqf.p, pl.end - pl.base, ctx->st);
            cudaEventRecord(freed[k], ctx->st);                   // the widening is the last reader of the raw chunk
            src = (const float2*)ctx->iqf.p;
        }
        rc = upload_items_range(ctx, off, len, nitems, b, e);
        if (rc == C8B_OK) rc = run_chunk(ctx, src, (const int64_t*)ctx->off.p, (const int32_t*)ctx->len.p, off, len, b, e, (c8b_frame*)ctx->frames.p,
                                         (uint8_t*)ctx->pdu.p, pdu_stride, pl.base, buf1[k]);
        if (!sc16) cudaEventRecord(freed[k], ctx->st);            // the front end is the last reader of the staged IQ
        if (rc == C8B_OK && er == cudaSuccess) {
            cudaStream_t sr = ctx->overlap ? ctx->stVit : ctx->st; // results follow the Viterbi pass of this chunk
            er = cudaMemcpyAsync(frames + b * maxf, (c8b_frame*)ctx->frames.p + b * maxf, (size_t)(e - b) * maxf * sizeof(c8b_frame),
                                 cudaMemcpyDeviceToHost, sr);
            if (er == cudaSuccess)
                er = cudaMemcpyAsync(pdu + b * maxf * pdu_stride, (uint8_t*)ctx->pdu.p + b * maxf * pdu_stride,
                                     (size_t)(e - b) * maxf * pdu_stride, cudaMemcpyDeviceToHost, sr);
        }
    }
    cudaError_t e2 = cudaStreamSynchronize(ctx->stCopy);
    const cudaError_t e3 = cudaStreamSynchronize(ctx->stVit), e4 = cudaStreamSynchronize(ctx->st);
    if (e2 == cudaSuccess) e2 = e3 != cudaSuccess ? e3 : e4;
    for (int k = 0; k < 2; k++) { cudaEventDestroy(copied[k]); cudaEventDestroy(freed[k]); }
    if (rc) return rc;
    if (er != cudaSuccess || e2 != cudaSuccess) { ctx->err = std::string("c8b_rx_batch: ") + cudaGetErrorString(er != cudaSuccess ? er : e2); return C8B_ERR_CUDA; }
    return C8B_OK;
}

int c8b_rx_batch(c8b_ctx* ctx, const float* h_iq, const int64_t* off, const int32_t* len, int nitems, c8b_frame* frames, uint8_t* pdu,
                 int64_t pdu_stride)
{
    return rx_batch_host(ctx, h_iq, nullptr, 0, off, len, nitems, frames, pdu, pdu_stride);
}

int c8b_rx_batch_sc16(c8b_ctx* ctx, const int16_t* h_iq, const int64_t* off, const int32_t* len, int nitems, c8b_frame* frames, uint8_t* pdu,
                      int64_t pdu_stride)
{
    return rx_batch_host(ctx, h_iq, nullptr, 1, off, len, nitems, frames, pdu, pdu_stride);
}

int c8b_rx_batch2(c8b_ctx* ctx, const float* h_iq0, const float* h_iq1, const int64_t* off, const int32_t* len, int nitems,
                  c8b_frame* frames, uint8_t* pdu, int64_t pdu_stride)
{
    if (!h_iq1) return C8B_ERR_ARG;
    return rx_batch_host(ctx, h_iq0, h_iq1, 0, off, len, nitems, frames, pdu, pdu_stride);
}

// ---- staged entry points (host buffers) -----------------------------------------------------------
int c8b_presiso(c8b_ctx* ctx, const float* h_iq, int64_t n, float* h_preac, float* h_preconj)
{
    if (!ctx || !h_iq || n < 0 || n > 0x7fffffff || !h_preac) return C8B_ERR_ARG;
    if (n == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    EN(iq, (size_t)n * sizeof(float2));
    EN(preac, (size_t)(n + 64) * sizeof(float));
    if (h_preconj) EN(preconj, (size_t)n * sizeof(float2));
    const int64_t off = 0; const int32_t len = (int32_t)n;
    int r = upload_items(ctx, &off, &len, 1);
    if (r) return r;
    CK(cudaMemcpyAsync(ctx->iq.p, h_iq, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, ctx->st));
    {
        StageTimer tm(ctx, C8B_K_PRESISO);
        c8b_launch_presiso((const float2*)ctx->iq.p, (const int64_t*)ctx->off.p, (const int32_t*)ctx->len.p, 1, len, 0, (float*)ctx->preac.p,
                           h_preconj ? (float2*)ctx->preconj.p : nullptr, nullptr, 0, ctx->st);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_preac, ctx->preac.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
    if (h_preconj) CK(cudaMemcpyAsync(h_preconj, ctx->preconj.p, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

int c8b_trigger(c8b_ctx* ctx, const float* h_preac, int64_t n, uint8_t* h_out)
{
    if (!ctx || !h_preac || n < 0 || !h_out) return C8B_ERR_ARG;
    if (n == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    EN(preac, (size_t)(n + 64) * sizeof(float));
    EN(trig, (size_t)n);
    CK(cudaMemcpyAsync(ctx->preac.p, h_preac, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->st));
    c8b_launch_trigger((const float*)ctx->preac.p, n, (uint8_t*)ctx->trig.p, ctx->st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_out, ctx->trig.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

static int stage_items(c8b_ctx* ctx, const float* h_iq, const int64_t* off, const int32_t* len, int n, ChunkPlan* pl)
{
    int r = check_items(ctx, off, len, n);
    if (r) return r;
    *pl = plan(off, len, n);
    if ((r = upload_items(ctx, off, len, n))) return r;
    EN(iq, (size_t)(pl->end - pl->base + 16) * sizeof(float2));
    CK(cudaMemcpyAsync(ctx->iq.p, reinterpret_cast<const float2*>(h_iq) + pl->base, (size_t)(pl->end - pl->base) * sizeof(float2),
                       cudaMemcpyHostToDevice, ctx->st));
    return C8B_OK;
}

int c8b_trigger_events(c8b_ctx* ctx, const float* h_preac, int64_t n, int from, int32_t* h_events, int cap, int* count, int* safe_end)
{
    if (!ctx || !h_preac || n <= 0 || n > 0x7ffffff0 || from < 0 || !h_events || cap <= 0 || !count) return C8B_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    EN(preac, (size_t)(n + 64) * sizeof(float));
    EN(mask, (size_t)((n + 31) / 32 + 2) * sizeof(uint32_t));
    EN(cand, c8b_detect_multi_scratch(1, (int)n));
    EN(scan, sizeof(c8b_scan));
    EN(trig, (size_t)(4 + 4 * cap) * sizeof(int32_t));
    const int64_t off = 0;
    const int32_t len = (int32_t)n;
    int r = upload_items(ctx, &off, &len, 1);
    if (r) return r;
    c8b_scan sc;
    memset(&sc, 0, sizeof sc);
    sc.from = from;
    CK(cudaMemcpyAsync(ctx->scan.p, &sc, sizeof sc, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->preac.p, h_preac, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->st));
    c8b_launch_trigger_events((const float*)ctx->preac.p, (int)n, (const int64_t*)ctx->off.p, (const int32_t*)ctx->len.p, (uint32_t*)ctx->mask.p,
                              (const c8b_scan*)ctx->scan.p, ctx->cand.p, cap, (int32_t*)ctx->trig.p, ctx->st);
    CK(cudaGetLastError());
    std::vector<int32_t> out((size_t)4 + 4 * cap);
    CK(cudaMemcpyAsync(out.data(), ctx->trig.p, out.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    *count = out[0];
    if (safe_end) *safe_end = out[1];
    for (int c = 0; c < out[0] && c < cap; c++) for (int k = 0; k < 4; k++) h_events[4 * c + k] = out[4 + 4 * c + k];
    return out[2] ? C8B_ERR_FULL : C8B_OK;
}

int c8b_detect(c8b_ctx* ctx, const float* h_iq, const int64_t* off, const int32_t* len, int nitems, c8b_frame* frames, float* h_chan)
{
    if (!ctx || !h_iq || !off || !len || nitems < 0 || !frames) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    if (nitems == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    ChunkPlan pl;
    if ((r = stage_items(ctx, h_iq, off, len, nitems, &pl))) return r;
    const int maxf = ctx->cfg.max_frames;
    const size_t ns = (size_t)nitems * maxf;
    const int maskStride = (pl.maxLen + 31) / 32 + 1;
    EN(preac, (size_t)(pl.end - pl.base + 64) * sizeof(float));
    EN(mask, (size_t)nitems * maskStride * sizeof(uint32_t));
    EN(frames, ns * sizeof(c8b_frame));
    EN(chan, ns * 64 * sizeof(float2));
    CK(cudaMemsetAsync(ctx->frames.p, 0, ns * sizeof(c8b_frame), ctx->st));
    CK(cudaMemsetAsync(ctx->chan.p, 0, ns * 64 * sizeof(float2), ctx->st));
    const float2* iq = (const float2*)ctx->iq.p - pl.base;
    {
        StageTimer tm(ctx, C8B_K_PRESISO);
        c8b_launch_presiso(iq, (const int64_t*)ctx->off.p, (const int32_t*)ctx->len.p, nitems, pl.maxLen, pl.base, (float*)ctx->preac.p, nullptr,
                           (uint32_t*)ctx->mask.p, maskStride, ctx->st);
    }
    {
        StageTimer tm(ctx, C8B_K_DETECT);
        (ctx->cfg.frontend_mode == 1 ? c8b_launch_detect : c8b_launch_detect_w)(
            ctx->d_lut, iq, (const int64_t*)ctx->off.p, (const int32_t*)ctx->len.p, nitems, 0, maxf, pl.base, (const float*)ctx->preac.p,
            (const uint32_t*)ctx->mask.p, maskStride, (c8b_frame*)ctx->frames.p, (float2*)ctx->chan.p, nullptr, ctx->st);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(frames, ctx->frames.p, ns * sizeof(c8b_frame), cudaMemcpyDeviceToHost, ctx->st));
    if (h_chan) CK(cudaMemcpyAsync(h_chan, ctx->chan.p, ns * 64 * sizeof(float2), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

int c8b_demod(c8b_ctx* ctx, const float* h_iq, const int64_t* off, const int32_t* len, int nitems, c8b_frame* frames, const float* h_chan,
              float* h_llr, int64_t llr_stride)
{
    if (!ctx || !h_iq || !off || !len || nitems < 0 || !frames || !h_chan || !h_llr || llr_stride <= 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    if (nitems == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    ChunkPlan pl;
    if ((r = stage_items(ctx, h_iq, off, len, nitems, &pl))) return r;
    const int maxf = ctx->cfg.max_frames;
    const size_t ns = (size_t)nitems * maxf;
    EN(frames, ns * sizeof(c8b_frame));
    EN(chan, ns * 64 * sizeof(float2));
    EN(hinv, ns * 64 * sizeof(float2));
    EN(llr, ns * llr_stride * sizeof(float));
    CK(cudaMemcpyAsync(ctx->frames.p, frames, ns * sizeof(c8b_frame), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->chan.p, h_chan, ns * 64 * sizeof(float2), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(ctx->llr.p, 0, ns * llr_stride * sizeof(float), ctx->st));
    const float2* iq = (const float2*)ctx->iq.p - pl.base;
    {
        StageTimer tm(ctx, C8B_K_HEADER);
        (ctx->cfg.frontend_mode == 1 ? c8b_launch_header : c8b_launch_header_w)(
            ctx->d_lut, iq, (const int64_t*)ctx->off.p, nitems, maxf, ctx->cfg.mupos, (c8b_frame*)ctx->frames.p, (const float2*)ctx->chan.p,
            (float2*)ctx->hinv.p, llr_stride, (float*)ctx->llr.p, ctx->st);
    }
    {
        StageTimer tm(ctx, C8B_K_DEMOD);
        c8b_launch_demod(ctx->d_lut, iq, (const int64_t*)ctx->off.p, nitems, maxf, (int)(llr_stride / 48 + 1), (const c8b_frame*)ctx->frames.p,
                         (const float2*)ctx->hinv.p, (float*)ctx->llr.p, ctx->st);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(frames, ctx->frames.p, ns * sizeof(c8b_frame), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(h_llr, ctx->llr.p, ns * llr_stride * sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

int c8b_demod2(c8b_ctx* ctx, const float* h_iq0, const float* h_iq1, const int64_t* off, const int32_t* len, int nitems, c8b_frame* frames,
               const float* h_chan, float* h_llr, int64_t llr_stride)
{
    if (!ctx || !h_iq0 || !h_iq1 || !off || !len || nitems < 0 || !frames || !h_chan || !h_llr || llr_stride <= 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    if (nitems == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    ChunkPlan pl;
    if ((r = stage_items(ctx, h_iq0, off, len, nitems, &pl))) return r;
    EN(iq1, (size_t)(pl.end - pl.base + 16) * sizeof(float2));
    CK(cudaMemcpyAsync(ctx->iq1.p, reinterpret_cast<const float2*>(h_iq1) + pl.base, (size_t)(pl.end - pl.base) * sizeof(float2),
                       cudaMemcpyHostToDevice, ctx->st));
    const int maxf = ctx->cfg.max_frames;
    const size_t ns = (size_t)nitems * maxf;
    EN(frames, ns * sizeof(c8b_frame));
    EN(chan, ns * 64 * sizeof(float2));
    EN(hinv, ns * 64 * sizeof(float2));
    EN(w2, ns * 264 * sizeof(float2));
    EN(llr, ns * llr_stride * sizeof(float));
    CK(cudaMemcpyAsync(ctx->frames.p, frames, ns * sizeof(c8b_frame), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->chan.p, h_chan, ns * 64 * sizeof(float2), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(ctx->llr.p, 0, ns * llr_stride * sizeof(float), ctx->st));
    const float2* iq = (const float2*)ctx->iq.p - pl.base;
    const float2* iq1 = (const float2*)ctx->iq1.p - pl.base;
    {
        StageTimer tm(ctx, C8B_K_HEADER);
        (ctx->cfg.frontend_mode == 1 ? c8b_launch_header2 : c8b_launch_header2_w)(
            ctx->d_lut, iq, iq1, (const int64_t*)ctx->off.p, nitems, maxf, ctx->cfg.mmse, (c8b_frame*)ctx->frames.p, (const float2*)ctx->chan.p,
            (float2*)ctx->hinv.p, (float2*)ctx->w2.p, llr_stride, ctx->st);
    }
    {
        StageTimer tm(ctx, C8B_K_DEMOD);
        const int maxSym = (int)(llr_stride / 48 + 1);
        c8b_launch_demod(ctx->d_lut, iq, (const int64_t*)ctx->off.p, nitems, maxf, maxSym, (const c8b_frame*)ctx->frames.p,
                         (const float2*)ctx->hinv.p, (float*)ctx->llr.p, ctx->st);
        c8b_launch_demod2(ctx->d_lut, iq, iq1, (const int64_t*)ctx->off.p, nitems, maxf, maxSym, (const c8b_frame*)ctx->frames.p,
                          (const float2*)ctx->w2.p, (float*)ctx->llr.p, ctx->st);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(frames, ctx->frames.p, ns * sizeof(c8b_frame), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(h_llr, ctx->llr.p, ns * llr_stride * sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

// ---- live stream ------------------------------------------------------------------------------------
// The reference's blocks run on an endless stream and keep their state between general_work calls.  Here the
// session keeps a device-resident window of the capture.  Every push appends samples and runs the whole chain on
// the window as ONE item; the detect kernel reports the latest point `safe` before which everything is decided
// and at which the trigger FSM is in its reset state (c8b_scan, lut.h).  Frames triggered before `safe` are
// returned; the window is then cut to [safe - 64, fill) -- 64 samples of presiso history -- and the next scan
// starts the FSM at `safe` with the signal block's consumed-until position carried over.  Because presiso is a
// pure function of the 64 previous samples and the FSM restarts from a state it provably had, the frames and
// PDUs are the ones a single pass over the whole capture finds, whatever the push sizes.
int c8b_stream_begin(c8b_ctx* ctx, int nant, int64_t window_samples)
{
    if (!ctx || (nant != 1 && nant != 2)) return C8B_ERR_ARG;
    if (window_samples <= 0) window_samples = (int64_t)1 << 22;
    if (window_samples < 4096 || window_samples > 0x7ffffff0) { ctx->err = "stream window: 4096 .. 2^31-16 samples"; return C8B_ERR_ARG; }
    if (ctx->cfg.max_frames < 2) { ctx->err = "c8b_stream_*: the ctx needs max_frames >= 2"; return C8B_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    for (int a = 0; a < nant; a++)
        for (int k = 0; k < 2; k++) {
            int r = ensure(ctx, ctx->sw[a][k], (size_t)(window_samples + 16) * sizeof(float2));
            if (r) return r;
        }
    EN(scan, sizeof(c8b_scan));
    ctx->strm = c8b_ctx::Stream();
    ctx->strm.open = true; ctx->strm.nant = nant; ctx->strm.cap = window_samples;
    return C8B_OK;
}

// one pass of the chain over the current window; appends the decided frames to the caller's arrays
static int stream_process(c8b_ctx* ctx, bool flush, c8b_frame* frames, int frames_cap, int* nframes, int64_t* frame_base, uint8_t* pdu,
                          int64_t pdu_stride, int* stalled)
{
    c8b_ctx::Stream& S = ctx->strm;
    *stalled = 0;
    if (S.fill <= S.from && !flush) return C8B_OK;                 // nothing new to scan
    if (S.fill == 0) return C8B_OK;
    const int maxf = ctx->cfg.max_frames;
    const int64_t off = 0;
    const int32_t len = (int32_t)S.fill;
    int r = upload_items(ctx, &off, &len, 1);
    if (r) return r;
    c8b_scan sc;
    memset(&sc, 0, sizeof sc);
    const int64_t p0 = S.posAbs - S.base;
    sc.from = S.from; sc.flush = flush ? 1 : 0;
    sc.pos0 = p0 < -0x40000000 ? -0x40000000 : (p0 > 0x7fffffff ? 0x7fffffff : (int32_t)p0);
    CK(cudaMemcpyAsync(ctx->scan.p, &sc, sizeof sc, cudaMemcpyHostToDevice, ctx->st));
    EN(frames, (size_t)maxf * sizeof(c8b_frame));
    EN(pdu, (size_t)maxf * pdu_stride);
    CK(cudaMemsetAsync(ctx->frames.p, 0, (size_t)maxf * sizeof(c8b_frame), ctx->st));
    ctx->scanDev = (c8b_scan*)ctx->scan.p;
    r = run_chunk(ctx, (const float2*)ctx->sw[0][S.cur].p, (const int64_t*)ctx->off.p, (const int32_t*)ctx->len.p, &off, &len, 0, 1,
                  (c8b_frame*)ctx->frames.p, (uint8_t*)ctx->pdu.p, pdu_stride, 0, S.nant == 2 ? (const float2*)ctx->sw[1][S.cur].p : nullptr);
    ctx->scanDev = nullptr;
    if (r) return r;
    join_viterbi(ctx);
    // read back through ONE pinned staging buffer: [frame records][scan result] in the first round trip, the PDU area of the
    // frames taken (contiguous slots) in the second -- not a copy per frame
    const size_t frBytes = (size_t)maxf * sizeof(c8b_frame), need = frBytes + 256 + (size_t)maxf * pdu_stride;
    if (ctx->hStageCap < need) {
        if (ctx->hStage) cudaFreeHost(ctx->hStage);
    if (ctx->oneFlagH) cudaFreeHost(ctx->oneFlagH);
    for (auto& sl : ctx->one) { if (sl.host) cudaFreeHost(sl.host); if (sl.dev.p) cudaFree(sl.dev.p); if (sl.ev) cudaEventDestroy(sl.ev); }
        ctx->hStage = nullptr; ctx->hStageCap = 0;
        CK(cudaHostAlloc(&ctx->hStage, need + need / 4, cudaHostAllocDefault));
        ctx->hStageCap = need + need / 4;
    }
    c8b_frame* fr = reinterpret_cast<c8b_frame*>(ctx->hStage);
    c8b_scan* hsc = reinterpret_cast<c8b_scan*>(reinterpret_cast<uint8_t*>(ctx->hStage) + frBytes);
    uint8_t* hpdu = reinterpret_cast<uint8_t*>(ctx->hStage) + frBytes + 256;
    CK(cudaMemcpyAsync(fr, ctx->frames.p, frBytes, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(hsc, ctx->scan.p, sizeof sc, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    sc = *hsc;
    *stalled = sc.stalled;
    int take = sc.nf;
    if (flush) {                                                   // end of stream: every record the batch semantics produced
        take = 0;
        while (take < maxf && fr[take].status != C8B_ST_EMPTY && fr[take].nsamp > 0) take++;
    }
    if (*nframes + take > frames_cap) { ctx->err = "c8b_stream_push: more frames than frames_cap"; return C8B_ERR_FULL; }
    bool anyPdu = false;
    for (int k = 0; k < take; k++) anyPdu |= fr[k].pdu_bytes > 0;
    if (anyPdu) {
        CK(cudaMemcpyAsync(hpdu, ctx->pdu.p, (size_t)take * pdu_stride, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    }
    for (int k = 0; k < take; k++) {
        const int o = (*nframes)++;
        frames[o] = fr[k];
        frames[o].item = o;
        frames[o].pdu_off = (int64_t)o * pdu_stride;
        frame_base[o] = S.base;
        if (fr[k].pdu_bytes > 0) memcpy(pdu + (size_t)o * pdu_stride, hpdu + (size_t)k * pdu_stride, (size_t)fr[k].pdu_bytes);
    }
    if (flush) {
        S.base += S.fill; S.fill = 0; S.from = 0; S.posAbs = S.base;
        return C8B_OK;
    }
    int64_t keep = sc.safe > 64 ? sc.safe - 64 : 0;                // 64 samples of presiso history before the restart point
    S.posAbs = S.base + sc.pos;
    if (keep == 0 && S.fill >= S.cap) {                            // a full window without one decidable point: drop it
        S.overruns++;
        S.base += S.fill; S.fill = 0; S.from = 0;
        if (S.posAbs < S.base) S.posAbs = S.base;
        return C8B_OK;
    }
    if (keep > 0) {
        const int64_t rest = S.fill - keep;
        for (int a = 0; a < S.nant; a++)
            CK(cudaMemcpyAsync(ctx->sw[a][S.cur ^ 1].p, (const float2*)ctx->sw[a][S.cur].p + keep, (size_t)rest * sizeof(float2),
                               cudaMemcpyDeviceToDevice, ctx->st));
        S.cur ^= 1; S.base += keep; S.fill = rest;
    }
    S.from = (int32_t)(sc.safe - keep);
    return C8B_OK;
}

int c8b_stream_push(c8b_ctx* ctx, const float* h_iq0, const float* h_iq1, int64_t n, int flush, c8b_frame* frames, int frames_cap,
                    int* nframes, int64_t* frame_base, uint8_t* pdu, int64_t pdu_stride)
{
    if (!ctx || n < 0 || (n > 0 && !h_iq0) || !frames || frames_cap < 0 || !nframes || !frame_base || !pdu || pdu_stride <= 0) return C8B_ERR_ARG;
    *nframes = 0;
    if (!ctx->strm.open) { ctx->err = "c8b_stream_push: no session (call c8b_stream_begin)"; return C8B_ERR_ARG; }
    c8b_ctx::Stream& S = ctx->strm;
    if (S.nant == 2 && n > 0 && !h_iq1) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    CK(cudaSetDevice(ctx->device));
    int64_t done = 0;
    for (;;) {
        const int64_t room = S.cap - S.fill;
        const int64_t take = n - done < room ? n - done : room;
        if (take > 0) {
            CK(cudaMemcpyAsync((float2*)ctx->sw[0][S.cur].p + S.fill, reinterpret_cast<const float2*>(h_iq0) + done, (size_t)take * sizeof(float2),
                               cudaMemcpyHostToDevice, ctx->st));
            if (S.nant == 2)
                CK(cudaMemcpyAsync((float2*)ctx->sw[1][S.cur].p + S.fill, reinterpret_cast<const float2*>(h_iq1) + done,
                                   (size_t)take * sizeof(float2), cudaMemcpyHostToDevice, ctx->st));
            S.fill += take; done += take;
        }
        const bool last = done == n;
        int stalled = 0;
        do {                                                       // frame records used up: the window holds more frames, go again
            const int64_t before = S.base + S.from;
            r = stream_process(ctx, false, frames, frames_cap, nframes, frame_base, pdu, pdu_stride, &stalled);
            if (r) return r;
            if (S.base + S.from == before) break;                  // no progress: wait for more samples
        } while (stalled == 2);
        if (!last) continue;
        if (flush) {
            r = stream_process(ctx, true, frames, frames_cap, nframes, frame_base, pdu, pdu_stride, &stalled);
            if (r) return r;
        }
        break;
    }
    return C8B_OK;
}

int c8b_stream_state(const c8b_ctx* ctx, int64_t* base, int64_t* fill, int64_t* overruns)
{
    if (!ctx || !ctx->strm.open) return C8B_ERR_ARG;
    if (base) *base = ctx->strm.base;
    if (fill) *fill = ctx->strm.fill;
    if (overruns) *overruns = ctx->strm.overruns;
    return C8B_OK;
}

// ---- transmit synthesiser ---------------------------------------------------------------------------
int c8b_tx_nsamp(int format, int mcs, int psdu_len)
{
    int nsym = 0, nslots = 0;
    if (!c8b_tx_geometry_host(format, mcs, psdu_len, &nsym, &nslots)) return C8B_ERR_ARG;
    return nslots * 80;
}

static int tx_run(c8b_ctx* ctx, const uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, float multiplier,
                  int seed, float* d_iq, float* d_iq1, int64_t iq_samples)
{
    if (seed < 1 || seed > 127) { ctx->err = "c8b_tx_batch: scrambler seed 1..127"; return C8B_ERR_ARG; }
    int maxSlots = 0;
    for (int i = 0; i < nframes; i++) {
        int nsym = 0, nslots = 0;
        const c8b_txframe& f = frames[i];
        if (!c8b_tx_geometry_host(f.format, f.mcs, f.psdu_len, &nsym, &nslots)) { ctx->err = "c8b_tx_batch: unsupported format / mcs / length"; return C8B_ERR_ARG; }
        if (c8b_tx_nss_host(f.format, f.mcs) == 2 && !d_iq1) { ctx->err = "c8b_tx_batch: a two-stream frame needs the two-antenna call (c8b_tx_batch2)"; return C8B_ERR_ARG; }
        if (f.psdu_off < 0 || f.psdu_off + f.psdu_len > psdu_bytes || f.out_off < 0 || f.out_off + (int64_t)nslots * 80 > iq_samples) {
            ctx->err = "c8b_tx_batch: frame outside the PSDU / IQ arena";
            return C8B_ERR_ARG;
        }
        if (nslots > maxSlots) maxSlots = nslots;
    }
    EN(txf, (size_t)nframes * sizeof(c8b_txframe));
    EN(txplan, c8b_tx_plan_bytes(nframes));
    CK(cudaMemcpyAsync(ctx->txf.p, frames, (size_t)nframes * sizeof(c8b_txframe), cudaMemcpyHostToDevice, ctx->st));
    uint32_t scr[4];
    c8b_tx_scrambler(seed, scr);
    c8b_launch_tx(ctx->d_lut, (const c8b_txframe*)ctx->txf.p, nframes, maxSlots, ctx->txplan.p, d_psdu, (float2*)d_iq, (float2*)d_iq1, multiplier, scr,
                  c8b_tx_eof_word(), ctx->st);
    CK(cudaGetLastError());
    return C8B_OK;
}

int c8b_tx_batch_dev(c8b_ctx* ctx, const uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, float multiplier,
                     int scrambler_seed, float* d_iq, int64_t iq_samples)
{
    return c8b_tx_batch2_dev(ctx, d_psdu, psdu_bytes, frames, nframes, multiplier, scrambler_seed, d_iq, nullptr, iq_samples);
}

int c8b_tx_batch2_dev(c8b_ctx* ctx, const uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, float multiplier,
                      int scrambler_seed, float* d_iq0, float* d_iq1, int64_t iq_samples)
{
    if (!ctx || !d_psdu || psdu_bytes < 0 || !frames || nframes < 0 || !d_iq0 || iq_samples < 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    if (nframes == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    r = tx_run(ctx, d_psdu, psdu_bytes, frames, nframes, multiplier, scrambler_seed, d_iq0, d_iq1, iq_samples);
    if (r) return r;
    CK(cudaStreamSynchronize(ctx->st));                            // the descriptor array is the caller's again
    return C8B_OK;
}

// ---- two-user MU-MIMO (genAmpduMu) ---------------------------------------------------------------------------------
int c8b_tx_mu_nsamp(int mcs0, int len0, int mcs1, int len1)
{
    int nsym = 0, nslots = 0;
    return c8b_tx_mu_geometry_host(mcs0, len0, mcs1, len1, &nsym, &nslots) ? nslots * 80 : C8B_ERR_ARG;
}

int c8b_tx_udp_parse_mu(const uint8_t* pkt, int pkt_len, c8b_txmu* f, const uint8_t** psdu0, const uint8_t** psdu1)
{
    if (!pkt || !f || pkt_len < 10 || pkt[0] != 3) return C8B_ERR_ARG;
    const int mcs0 = pkt[1], nss0 = pkt[2], len0 = pkt[3] | (pkt[4] << 8), mcs1 = pkt[5], nss1 = pkt[6], len1 = pkt[7] | (pkt[8] << 8), gid = pkt[9];
    if (len0 + len1 > 4095 || pkt_len < len0 + len1 + 10) return C8B_ERR_ARG;     // pktPop :127-130
    if (nss0 != 1 || nss1 != 1 || gid < 1 || gid > 62 || c8b_tx_mu_nsamp(mcs0, len0, mcs1, len1) < 0) return C8B_ERR_ARG;
    memset(f, 0, sizeof(*f));
    f->mcs[0] = mcs0; f->mcs[1] = mcs1; f->psdu_len[0] = len0; f->psdu_len[1] = len1; f->group_id = gid;
    if (psdu0) *psdu0 = pkt + 10;
    if (psdu1) *psdu1 = pkt + 10 + len0;
    return C8B_OK;
}

int c8b_tx_udp_parse_bfq(const uint8_t* pkt, int pkt_len, float* q_out)
{
    if (!pkt || !q_out || pkt_len != 2049 || pkt[0] != 10) return C8B_ERR_ARG;  // C8P_F_VHT_BFQ, exactly 256 complex floats
    memcpy(q_out, pkt + 1, 2048);
    return C8B_OK;
}

static int tx_run_mu(c8b_ctx* ctx, const uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txmu* frames, int nframes, const float* d_q, int nq,
                     float multiplier, int seed, float* d_iq0, float* d_iq1, int64_t iq_samples)
{
    if (seed < 1 || seed > 127) { ctx->err = "c8b_tx_mu_batch: scrambler seed 1..127"; return C8B_ERR_ARG; }
    int maxSlots = 0;
    for (int i = 0; i < nframes; i++) {
        int nsym = 0, nslots = 0;
        const c8b_txmu& f = frames[i];
        if (!c8b_tx_mu_geometry_host(f.mcs[0], f.psdu_len[0], f.mcs[1], f.psdu_len[1], &nsym, &nslots) || f.group_id < 1 || f.group_id > 62) {
            ctx->err = "c8b_tx_mu_batch: VHT MCS 0-8, A-MPDUs of 4..4092 bytes (multiples of 4), group id 1..62";
            return C8B_ERR_ARG;
        }
        for (int u = 0; u < 2; u++)
            if (f.psdu_off[u] < 0 || f.psdu_off[u] + f.psdu_len[u] > psdu_bytes) { ctx->err = "c8b_tx_mu_batch: A-MPDU outside the PSDU arena"; return C8B_ERR_ARG; }
        if (f.out_off < 0 || f.out_off + (int64_t)nslots * 80 > iq_samples || f.q_index < 0 || f.q_index >= nq) {
            ctx->err = "c8b_tx_mu_batch: frame outside the IQ arena / q_index outside the matrix sets";
            return C8B_ERR_ARG;
        }
        if (nslots > maxSlots) maxSlots = nslots;
    }
    EN(txf, (size_t)nframes * sizeof(c8b_txmu));
    EN(txplan, c8b_tx_plan_bytes(2 * nframes));
    CK(cudaMemcpyAsync(ctx->txf.p, frames, (size_t)nframes * sizeof(c8b_txmu), cudaMemcpyHostToDevice, ctx->st));
    uint32_t scr[4];
    c8b_tx_scrambler(seed, scr);
    c8b_launch_tx_mu(ctx->d_lut, (const c8b_txmu*)ctx->txf.p, nframes, maxSlots, ctx->txplan.p, d_psdu, (const float2*)d_q, (float2*)d_iq0, (float2*)d_iq1,
                     multiplier, scr, c8b_tx_eof_word(), ctx->st);
    CK(cudaGetLastError());
    return C8B_OK;
}

int c8b_tx_mu_batch_dev(c8b_ctx* ctx, const uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txmu* frames, int nframes, const float* d_q, int nq,
                        float multiplier, int scrambler_seed, float* d_iq0, float* d_iq1, int64_t iq_samples)
{
    if (!ctx || !d_psdu || psdu_bytes < 0 || !frames || nframes < 0 || !d_q || nq < 1 || !d_iq0 || !d_iq1 || iq_samples < 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    if (nframes == 0) return C8B_OK;
    CK(cudaSetDevice(ctx->device));
    r = tx_run_mu(ctx, d_psdu, psdu_bytes, frames, nframes, d_q, nq, multiplier, scrambler_seed, d_iq0, d_iq1, iq_samples);
    if (r) return r;
    CK(cudaStreamSynchronize(ctx->st));                            // the descriptor array is the caller's again
    return C8B_OK;
}

int c8b_tx_mu_batch(c8b_ctx* ctx, const uint8_t* h_psdu, int64_t psdu_bytes, const c8b_txmu* frames, int nframes, const float* h_q, int nq,
                    float multiplier, int scrambler_seed, float* h_iq0, float* h_iq1, int64_t iq_samples)
{
    if (!ctx || !h_psdu || psdu_bytes < 0 || !frames || nframes < 0 || !h_q || nq < 1 || !h_iq0 || !h_iq1 || iq_samples < 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    CK(cudaSetDevice(ctx->device));
    const size_t arena = ((size_t)(iq_samples + 16) * sizeof(float2) + 255) & ~(size_t)255;
    const size_t qBytes = (size_t)nq * 256 * sizeof(float2), qOff = ((size_t)psdu_bytes + 16 + 255) & ~(size_t)255;
    EN(txpsdu, qOff + qBytes);                                      // [A-MPDUs][matrix sets]
    EN(txiq, arena * 2);
    float* d0 = (float*)ctx->txiq.p;
    float* d1 = (float*)((uint8_t*)ctx->txiq.p + arena);
    const float* dq = (const float*)((uint8_t*)ctx->txpsdu.p + qOff);
    CK(cudaMemcpyAsync(ctx->txpsdu.p, h_psdu, (size_t)psdu_bytes, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync((void*)dq, h_q, qBytes, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(ctx->txiq.p, 0, arena * 2, ctx->st));
    if (nframes > 0) {
        r = tx_run_mu(ctx, (const uint8_t*)ctx->txpsdu.p, psdu_bytes, frames, nframes, dq, nq, multiplier, scrambler_seed, d0, d1, iq_samples);
        if (r) return r;
    }
    CK(cudaMemcpyAsync(h_iq0, d0, (size_t)iq_samples * sizeof(float2), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(h_iq1, d1, (size_t)iq_samples * sizeof(float2), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

int c8b_tx_random_psdu_dev(c8b_ctx* ctx, uint8_t* d_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, uint64_t seed)
{
    if (!ctx || !d_psdu || psdu_bytes < 0 || !frames || nframes < 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    if (nframes == 0) return C8B_OK;
    for (int i = 0; i < nframes; i++) {
        const c8b_txframe& f = frames[i];
        const bool vht = f.format == C8B_F_VHT;
        if (f.psdu_off < 0 || f.psdu_len < (vht ? 12 : 5) || f.psdu_len > 4095 || (vht && (f.psdu_len & 3)) || f.psdu_off + f.psdu_len > psdu_bytes) {
            ctx->err = "c8b_tx_random_psdu_dev: PSDU region (VHT: 12.. bytes, multiple of 4; else 5.. bytes) outside the arena";
            return C8B_ERR_ARG;
        }
    }
    CK(cudaSetDevice(ctx->device));
    EN(txf, (size_t)nframes * sizeof(c8b_txframe));
    CK(cudaMemcpyAsync(ctx->txf.p, frames, (size_t)nframes * sizeof(c8b_txframe), cudaMemcpyHostToDevice, ctx->st));
    c8b_launch_tx_fill(ctx->d_lut, (const c8b_txframe*)ctx->txf.p, nframes, d_psdu, seed, ctx->st);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

int c8b_tx_batch(c8b_ctx* ctx, const uint8_t* h_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, float multiplier,
                 int scrambler_seed, float* h_iq, int64_t iq_samples)
{
    return c8b_tx_batch2(ctx, h_psdu, psdu_bytes, frames, nframes, multiplier, scrambler_seed, h_iq, nullptr, iq_samples);
}

int c8b_tx_batch2(c8b_ctx* ctx, const uint8_t* h_psdu, int64_t psdu_bytes, const c8b_txframe* frames, int nframes, float multiplier,
                  int scrambler_seed, float* h_iq0, float* h_iq1, int64_t iq_samples)
{
    if (!ctx || !h_psdu || psdu_bytes < 0 || !frames || nframes < 0 || !h_iq0 || iq_samples < 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    CK(cudaSetDevice(ctx->device));
    const size_t arena = ((size_t)(iq_samples + 16) * sizeof(float2) + 255) & ~(size_t)255;
    EN(txpsdu, (size_t)psdu_bytes + 16);
    EN(txiq, arena * (h_iq1 ? 2 : 1));
    float* d0 = (float*)ctx->txiq.p;
    float* d1 = h_iq1 ? (float*)((uint8_t*)ctx->txiq.p + arena) : nullptr;
    CK(cudaMemcpyAsync(ctx->txpsdu.p, h_psdu, (size_t)psdu_bytes, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(ctx->txiq.p, 0, arena * (h_iq1 ? 2 : 1), ctx->st));
    if (nframes > 0) {
        r = tx_run(ctx, (const uint8_t*)ctx->txpsdu.p, psdu_bytes, frames, nframes, multiplier, scrambler_seed, d0, d1, iq_samples);
        if (r) return r;
    }
    CK(cudaMemcpyAsync(h_iq0, d0, (size_t)iq_samples * sizeof(float2), cudaMemcpyDeviceToHost, ctx->st));
    if (h_iq1) CK(cudaMemcpyAsync(h_iq1, d1, (size_t)iq_samples * sizeof(float2), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return C8B_OK;
}

// ---- MAC -> PHY UDP framing (lib/pktgen_impl.cc:57-70,96-118; tools/phy80211.py:1126-1137) -------------------------
int c8b_tx_udp_parse(const uint8_t* pkt, int pkt_len, c8b_txframe* f, const uint8_t** psdu)
{
    if (!pkt || !f || pkt_len < 5) return C8B_ERR_ARG;                       // msgRead drops datagrams under 5 bytes (:65-67)
    const int format = pkt[0], mcs = pkt[1], nss = pkt[2], len = pkt[3] | (pkt[4] << 8);
    if (format == 3) return C8B_ERR_ARG;                                      // C8P_F_VHT_MU: two users per datagram, not synthesised here
    if (len > 4095 || pkt_len < len + 5) return C8B_ERR_ARG;                  // pktPop :108-111
    if (nss != 1 && nss != 2) return C8B_ERR_ARG;
    int code = mcs;                                                           // descriptor mcs: HT 8-15 already say two streams, VHT adds 16
    if (format == 2 && nss == 2) code = mcs + 16;
    if (format == 0 && nss != 1) return C8B_ERR_ARG;
    if (format == 1 && (nss == 2) != (mcs >= 8)) return C8B_ERR_ARG;
    if (c8b_tx_nsamp(format, code, len) < 0) return C8B_ERR_ARG;
    memset(f, 0, sizeof(*f));
    f->format = format; f->mcs = code; f->psdu_len = len;
    if (psdu) *psdu = pkt + 5;
    return nss;
}

int c8b_tx_from_udp(c8b_ctx* ctx, const uint8_t* pkts, const int64_t* pkt_off, const int32_t* pkt_len, int npkts, int gap, float multiplier,
                    int scrambler_seed, float* h_iq, int64_t iq_cap, int64_t* iq_used, c8b_txframe* frames_out)
{
    if (!ctx || !pkts || !pkt_off || !pkt_len || npkts < 0 || gap < 0 || !h_iq || iq_cap < 0) return C8B_ERR_ARG;
    std::vector<c8b_txframe> fr;
    std::vector<uint8_t> arena;
    int64_t pos = 0;
    for (int k = 0; k < npkts; k++) {
        c8b_txframe f;
        const uint8_t* body = nullptr;
        if (frames_out) { memset(&frames_out[k], 0, sizeof(c8b_txframe)); frames_out[k].psdu_len = -1; }
        if (pkt_off[k] < 0 || pkt_len[k] < 0) return C8B_ERR_ARG;
        if (c8b_tx_udp_parse(pkts + pkt_off[k], pkt_len[k], &f, &body) != 1) continue;     // (two-stream frames need two arenas: c8b_tx_batch2)
        f.psdu_off = (int64_t)arena.size();
        arena.insert(arena.end(), body, body + f.psdu_len);
        while (arena.size() & 3) arena.push_back(0);
        f.out_off = pos + gap;
        pos = f.out_off + c8b_tx_nsamp(f.format, f.mcs, f.psdu_len);
        if (frames_out) frames_out[k] = f;
        fr.push_back(f);
    }
    pos += gap;
    if (pos > iq_cap) { ctx->err = "c8b_tx_from_udp: iq_cap too small"; return C8B_ERR_FULL; }
    if (iq_used) *iq_used = pos;
    arena.resize(arena.size() + 16, 0);
    const int rc = c8b_tx_batch(ctx, arena.data(), (int64_t)arena.size(), fr.data(), (int)fr.size(), multiplier, scrambler_seed, h_iq, pos);
    return rc ? rc : (int)fr.size();
}

}  // extern "C"

// ---- one frame per call: what the demod / demod2 / decode blocks of csrc/blocks.cu run ---------------------------------
// c8b_demod / c8b_decode stage every array with its own copy and clear their scratch (they serve arbitrary batches); a
// gr::block call hands over ONE frame, and what it costs is the number of driver calls.  Here the inputs of the frame
// travel as one packed block (one pinned staging buffer, one H2D copy), the kernels run on it in place, and the results
// come back as one block (one D2H copy, one synchronise).
namespace {
struct OneHdr {                      // head of the packed block, identical on host and device
    int64_t off;                     // item table of the one item: offset 0 ...
    int32_t len, pad;                // ... and its length in samples
    c8b_frame f;                     // the frame record (in: detect fields; out: everything)
    float chan[128];                 // legacy channel (tag "chan")
};
constexpr size_t ONE_HDR = (sizeof(OneHdr) + 255) & ~(size_t)255;

int one_buffers(c8b_ctx* ctx, int slot, size_t bytes)
{
    c8b_ctx::OneSlot& sl = ctx->one[slot];
    if (sl.hostCap < bytes) {
        if (sl.host) cudaFreeHost(sl.host);
        sl.host = nullptr; sl.hostCap = 0;
        const size_t want = bytes + bytes / 2 + 65536;
        if (cudaHostAlloc(&sl.host, want, cudaHostAllocDefault) != cudaSuccess) { ctx->err = "cudaHostAlloc (per-frame staging)"; return C8B_ERR_NOMEM; }
        sl.hostCap = want;
    }
    if (!sl.ev && cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming) != cudaSuccess) { ctx->err = "cudaEventCreate"; return C8B_ERR_CUDA; }
    if (!ctx->oneFlagH) {
        if (cudaHostAlloc((void**)&ctx->oneFlagH, 256, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void**)&ctx->oneFlagD, ctx->oneFlagH, 0) != cudaSuccess) { ctx->err = "cudaHostAlloc (completion flags)"; return C8B_ERR_NOMEM; }
        memset(ctx->oneFlagH, 0, 256);
    }
    return ensure(ctx, sl.dev, bytes + 256);
}
}  // namespace

// header states + per-symbol demod of one frame, asynchronous (slot 0 / 1, collected in submission order).  iq0 / iq1: n
// samples per antenna (the stream the signal block hands on, behind the 224 samples it consumed); f: detect fields set;
// chan: 64 complex.  Layout of a slot: [soft bits][header][samples][1/H, weights] -- ONE H2D copy (header + samples), ONE
// D2H copy (soft bits + header).  collect: *f = the finished record, *llr_out = its soft bits in the slot's pinned buffer
// (llr_n floats, valid until the slot is submitted again).
int c8b_one_demod_submit(c8b_ctx* ctx, int slot, int nant, const float* iq0, const float* iq1, int n, const c8b_frame* f, const float* chan)
{
    if (!ctx || slot < 0 || slot >= 2 || !iq0 || (nant == 2 && !iq1) || n <= 0 || !f || !chan) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    CK(cudaSetDevice(ctx->device));
    const int64_t stride = llr_stride_for(f->nsamp > 0 ? f->nsamp : n) * (nant == 2 ? 2 : 1);
    const size_t llrBytes = ((size_t)stride * sizeof(float) + 255) & ~(size_t)255, iqBytes = ((size_t)n * sizeof(float2) + 255) & ~(size_t)255;
    const size_t oHdr = llrBytes, oIq0 = oHdr + ONE_HDR, oIq1 = oIq0 + iqBytes, oAux = oIq1 + (nant == 2 ? iqBytes : 0);
    const size_t total = oAux + 64 * sizeof(float2) + 264 * sizeof(float2) + 256;
    if ((r = one_buffers(ctx, slot, total))) return r;
    c8b_ctx::OneSlot& sl = ctx->one[slot];
    uint8_t* h = reinterpret_cast<uint8_t*>(sl.host);
    uint8_t* d = reinterpret_cast<uint8_t*>(sl.dev.p);
    OneHdr* hh = reinterpret_cast<OneHdr*>(h + oHdr);
    hh->off = 0; hh->len = n; hh->pad = 0; hh->f = *f;
    memcpy(hh->chan, chan, sizeof(hh->chan));
    memcpy(h + oIq0, iq0, (size_t)n * sizeof(float2));
    if (nant == 2) memcpy(h + oIq1, iq1, (size_t)n * sizeof(float2));
    CK(cudaMemcpyAsync(d + oHdr, h + oHdr, oAux - oHdr, cudaMemcpyHostToDevice, ctx->st));
    OneHdr* dh = reinterpret_cast<OneHdr*>(d + oHdr);
    const float2* dq0 = reinterpret_cast<const float2*>(d + oIq0);
    const float2* dq1 = nant == 2 ? reinterpret_cast<const float2*>(d + oIq1) : nullptr;
    float2* dhinv = reinterpret_cast<float2*>(d + oAux);
    float2* dw2 = dhinv + 64;
    float* dllr = reinterpret_cast<float*>(d);
    {
        StageTimer tm(ctx, C8B_K_HEADER);
        if (nant == 2)
            (ctx->cfg.frontend_mode == 1 ? c8b_launch_header2 : c8b_launch_header2_w)(ctx->d_lut, dq0, dq1, &dh->off, 1, 1, ctx->cfg.mmse, &dh->f,
                                                                                       reinterpret_cast<const float2*>(dh->chan), dhinv, dw2, stride, ctx->st);
        else
            (ctx->cfg.frontend_mode == 1 ? c8b_launch_header : c8b_launch_header_w)(ctx->d_lut, dq0, &dh->off, 1, 1, ctx->cfg.mupos, &dh->f,
                                                                                     reinterpret_cast<const float2*>(dh->chan), dhinv, stride, dllr, ctx->st);
    }
    {
        StageTimer tm(ctx, C8B_K_DEMOD);
        const int maxSym = (int)(stride / 48 + 1);
        c8b_launch_demod(ctx->d_lut, dq0, &dh->off, 1, 1, maxSym, &dh->f, dhinv, dllr, ctx->st);
        if (nant == 2) c8b_launch_demod2(ctx->d_lut, dq0, dq1, &dh->off, 1, 1, maxSym, &dh->f, dw2, dllr, ctx->st);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h, d, oHdr + ONE_HDR, cudaMemcpyDeviceToHost, ctx->st));       // the soft bits and the frame record in one copy
    sl.seq = ++ctx->oneSeq;
    sl.aux = (int64_t)oHdr;
    sl.aux2 = stride;
    c8b_launch_flag(ctx->oneFlagD + 8 + slot, sl.seq, ctx->st);
    CK(cudaEventRecord(sl.ev, ctx->st));
    return C8B_OK;
}

int c8b_one_demod_collect(c8b_ctx* ctx, int slot, int wait, c8b_frame* f, const float** llr_out, int* llr_n)
{
    if (!ctx || slot < 0 || slot >= 2 || !ctx->one[slot].ev || !f || !llr_out || !llr_n) return C8B_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    c8b_ctx::OneSlot& sl = ctx->one[slot];
    const volatile uint32_t* flag = ctx->oneFlagH + 8 + slot;
    if (wait) {
        if (c8b_wait_flag(flag, sl.seq, ctx->st) != 0) CK(cudaEventSynchronize(sl.ev));
    } else if (*flag != sl.seq) return 0;
    const uint8_t* h = reinterpret_cast<const uint8_t*>(sl.host);
    *f = reinterpret_cast<const OneHdr*>(h + sl.aux)->f;
    *llr_out = reinterpret_cast<const float*>(h);
    *llr_n = (int)sl.aux2;
    return 1;
}

// decode of one frame, asynchronous: submit copies the frame record and its f->total soft bits into staging slot `slot`
// (0 .. C8B_ONE_SLOTS-1; the caller cycles through them and collects in the same order) and enqueues copy-in, the decode
// kernel and copy-out on the ctx stream without waiting; collect(wait = 0) says whether that slot has finished (1) or not
// (0), collect(wait = 1) waits for it.  *pdu_out: the PDU records (f->pdu_bytes bytes) in the slot's pinned buffer, valid
// until the slot is submitted again.  Layout of a slot: [PDU area][header][soft bits] -- one H2D copy (header + soft bits),
// one D2H copy (PDU area + header).
int c8b_one_decode_submit(c8b_ctx* ctx, int slot, const c8b_frame* f, const float* llr, int nllr)
{
    if (!ctx || slot < 0 || slot >= C8B_ONE_SLOTS || !f || !llr || nllr < 0) return C8B_ERR_ARG;
    int r = need_lut(ctx);
    if (r) return r;
    CK(cudaSetDevice(ctx->device));
    const size_t llrBytes = (size_t)nllr * sizeof(float), pduCap = 2 * 4400 + 256;
    const size_t oHdr = pduCap, oLlr = oHdr + ONE_HDR, total = oLlr + llrBytes + 256;
    if ((r = one_buffers(ctx, slot, total))) return r;
    if ((r = ensure_surv(ctx))) return r;
    uint8_t* h = reinterpret_cast<uint8_t*>(ctx->one[slot].host);
    uint8_t* d = reinterpret_cast<uint8_t*>(ctx->one[slot].dev.p);
    OneHdr* hh = reinterpret_cast<OneHdr*>(h + oHdr);
    hh->off = 0; hh->len = 0; hh->pad = 0; hh->f = *f;
    hh->f.llr_off = 0;
    memcpy(h + oLlr, llr, llrBytes);
    CK(cudaMemcpyAsync(d + oHdr, h + oHdr, ONE_HDR + llrBytes, cudaMemcpyHostToDevice, ctx->st));
    OneHdr* dh = reinterpret_cast<OneHdr*>(d + oHdr);
    {
        StageTimer tm(ctx, C8B_K_VITERBI);
        c8b_launch_viterbi(ctx->d_lut, &dh->f, 1, reinterpret_cast<const float*>(d + oLlr), nllr, (uint2*)ctx->surv.p, ctx->survWarps, d, (int64_t)(2 * 4400),
                           nullptr, 0, ctx->d_counter, 1, ctx->st);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h, d, oHdr + ONE_HDR, cudaMemcpyDeviceToHost, ctx->st));
    ctx->one[slot].seq = ++ctx->oneSeq;
    c8b_launch_flag(ctx->oneFlagD + 8 + slot, ctx->one[slot].seq, ctx->st);          // (word 0 belongs to c8b_one_demod)
    CK(cudaEventRecord(ctx->one[slot].ev, ctx->st));
    return C8B_OK;
}

int c8b_one_decode_collect(c8b_ctx* ctx, int slot, int wait, c8b_frame* f, const uint8_t** pdu_out)
{
    if (!ctx || slot < 0 || slot >= C8B_ONE_SLOTS || !ctx->one[slot].ev || !f || !pdu_out) return C8B_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const volatile uint32_t* flag = ctx->oneFlagH + 8 + slot;
    if (wait) {
        if (c8b_wait_flag(flag, ctx->one[slot].seq, ctx->st) != 0) CK(cudaEventSynchronize(ctx->one[slot].ev));
    } else if (*flag != ctx->one[slot].seq) return 0;
    const uint8_t* h = reinterpret_cast<const uint8_t*>(ctx->one[slot].host);
    *f = reinterpret_cast<const OneHdr*>(h + 2 * 4400 + 256)->f;
    *pdu_out = h;
    return 1;
}
