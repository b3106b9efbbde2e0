// TEST INFRASTRUCTURE: plays the GNU Radio scheduler for the compiled block shells (gr-ieee80211_b200/gr/lib/
// rx_blocks_impl.cc, built against the miniature runtime of tests/gr_mock/include) -- and, linked as build/run_chain_ref
// against the reference's own unmodified *_impl.cc objects (oracle/_ref, no GPU involved), for the reference's blocks, so
// the two can be compared dump against dump: wires trigger -> sync -> signal[2] ->
// demod[2] -> decode like examples/rx.grc / rx2.grc, feeds presiso's outputs and the capture from files, calls
// general_work() with pseudo-random sizes until nothing moves, and dumps the published messages and all stream tags.
//   run_chain NANT MUPOS MUGID SEED MAXCALL IFDEBUG INDIR OUTFILE
// INDIR holds preac.f32, preconj.c64, sig.c64 [, sig1.c64]; OUTFILE gets lines
//   MSG <hex bytes>          one per message on decode's port "out", in order
//   TAG <block> <offset> <key> <value>
#include <gnuradio/ieee80211/decode.h>
#include <gnuradio/ieee80211/demod.h>
#include <gnuradio/ieee80211/demod2.h>
#include <gnuradio/ieee80211/signal.h>
#include <gnuradio/ieee80211/signal2.h>
#include <gnuradio/ieee80211/sync.h>
#include <gnuradio/ieee80211/trigger.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

using gr::mock::edge;

static std::vector<char> slurp(const std::string& p)
{
    std::ifstream f(p, std::ios::binary);
    if (!f) { std::cerr << "cannot read " << p << std::endl; exit(2); }
    return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
static std::map<std::string, std::pair<double, long>> g_wall;   // block name -> (ms inside general_work, calls)
static std::mutex g_wall_mu;
static uint32_t rng_state = 1;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

struct Graph {
    std::vector<std::unique_ptr<edge>> edges;
    edge* make_edge(int item) { edges.emplace_back(new edge); edges.back()->item = item; return edges.back().get(); }
    void connect(gr::block& a, int pa, gr::block& b, int pb)
    {
        edge* e = make_edge(a.output_signature()->sizeof_stream_item(pa));
        a.mock_out.at(pa).push_back(e);
        b.mock_in.at(pb) = e;
    }
    edge* source(const std::vector<char>& data, int item, gr::block& b, int pb)
    {
        edge* e = make_edge(item);
        e->data = data; e->nwritten = data.size() / item;
        b.mock_in.at(pb) = e;
        return e;
    }
};

// one general_work() call with the sizes a scheduler could pick; returns whether anything moved
static bool call(gr::block& b, int maxCall, bool big)
{
    const int nin = (int)b.mock_in.size(), nout = (int)b.mock_out.size();
    int cap = big ? maxCall : 1 + (int)(rnd() % (uint32_t)maxCall);
    size_t minAvail = (size_t)-1;
    for (edge* e : b.mock_in) minAvail = std::min(minAvail, e->avail());
    int noutput = cap;
    gr_vector_int req(nin), ninput(nin);
    if (nout > 0) {
        noutput = (int)std::min<size_t>((size_t)cap, minAvail);
        if (noutput == 0 && (b.name() == "demod" || b.name() == "demod2")) noutput = cap;   // output space alone wakes a block too
        // end of a finite capture: signal / demod size their work from ninput_items, so the closing rounds offer them the whole
        // output space (the reference's demod needs noutput > nCBPS to finish the last symbols; oracle/ref_chain.cc does the same)
        if (big && b.name() != "trigger" && b.name() != "sync") noutput = cap;
    }
    b.forecast(noutput, req);
    const int extra = big ? 0 : (int)(rnd() % 64u);
    for (int k = 0; k < nin; k++) ninput[k] = (int)std::min<size_t>(b.mock_in[k]->avail(), (size_t)std::max(req[k], noutput) + extra);
    int most = 0;
    for (int v : ninput) most = std::max(most, v);
    if (noutput <= 0 && most <= 0) return false;
    gr_vector_const_void_star in(nin);
    for (int k = 0; k < nin; k++) in[k] = b.mock_in[k]->data.data() + b.mock_in[k]->head;
    std::vector<std::vector<char>> obuf(nout);
    gr_vector_void_star out(nout);
    for (int k = 0; k < nout; k++) {
        obuf[k].assign((size_t)std::max(noutput, 1) * b.output_signature()->sizeof_stream_item(k), 0);
        out[k] = obuf[k].data();
    }
    b.mock_consumed = 0;
    const size_t tagsBefore = b.mock_tags_added.size(), msgBefore = b.mock_messages.size();
    const auto tw0 = std::chrono::steady_clock::now();
    const int produced = b.general_work(noutput, ninput, in, out);
    {
        std::lock_guard<std::mutex> lk(g_wall_mu);
        auto& acc = g_wall[b.name()];
        acc.first += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
        acc.second++;
    }
    if (produced < 0 || produced > noutput || b.mock_consumed < 0) { std::cerr << b.name() << ": bad accounting" << std::endl; exit(3); }
    for (int k = 0; k < nin; k++) {
        edge* e = b.mock_in[k];
        if ((size_t)b.mock_consumed > e->avail()) { std::cerr << b.name() << ": consumed more than available" << std::endl; exit(3); }
        e->head += (size_t)b.mock_consumed * e->item;              // (no erase per call: the sources hold the whole capture)
        e->nread += b.mock_consumed;
        if (e->head > (1u << 22) && e->head * 2 > e->data.size()) { e->data.erase(e->data.begin(), e->data.begin() + (ptrdiff_t)e->head); e->head = 0; }
        if (e->tags.size() > 64) {
            std::vector<gr::tag_t> keep;
            for (auto& t : e->tags) if (t.offset >= e->nread) keep.push_back(t);
            e->tags.swap(keep);
        }
    }
    for (int k = 0; k < nout; k++) {
        const size_t bytes = (size_t)produced * b.output_signature()->sizeof_stream_item(k);
        for (edge* e : b.mock_out[k]) { e->data.insert(e->data.end(), obuf[k].begin(), obuf[k].begin() + bytes); e->nwritten += produced; }
        b.mock_written[k] += produced;
    }
    return b.mock_consumed > 0 || produced > 0 || b.mock_tags_added.size() != tagsBefore || b.mock_messages.size() != msgBefore;
}

// ---- thread-per-block mode (RUN_CHAIN_TPB=1): GNU Radio 3.10's default scheduler gives every block its own thread; the
// blocks of a chain then work on different frames at the same time and the chain runs at the pace of its slowest block.
// One mutex guards the edges; a call snapshots its input windows (items + tags) under it, runs general_work on the
// snapshot with the lock released, and commits consumption / output / tags under it again.  Call sizes are the largest
// the edges allow (what the executor offers a block that keeps up), capped at MAXCALL.
struct Tpb {
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> epoch{ 0 };
    std::atomic<long long> lastNs{ 0 };
    std::atomic<bool> stop{ false }, flush{ false };
    std::chrono::steady_clock::time_point t0;
};

static bool call_tpb(gr::block& b, const std::vector<edge*>& rin, const std::vector<std::vector<edge*>>& rout, int maxCall, Tpb& T)
{
    const int nin = (int)rin.size(), nout = (int)rout.size();
    std::vector<edge> pin((size_t)nin), pout((size_t)nout);
    std::vector<const char*> inp((size_t)nin, nullptr);
    gr_vector_int req(nin), ninput(nin);
    int noutput = maxCall;
    const bool flush = T.flush.load();
    {
        std::unique_lock<std::mutex> lk(T.mu);
        size_t minAvail = (size_t)-1;
        for (edge* e : rin) minAvail = std::min(minAvail, e->avail());
        if (nout > 0) {
            noutput = (int)std::min<size_t>((size_t)maxCall, minAvail);
            if (noutput == 0 && (b.name() == "demod" || b.name() == "demod2")) noutput = maxCall;
            if (flush && b.name() != "trigger" && b.name() != "sync") noutput = maxCall;
        }
        b.forecast(noutput, req);
        int most = 0;
        for (int k = 0; k < nin; k++) { ninput[k] = (int)std::min<size_t>(rin[k]->avail(), (size_t)std::max(req[k], noutput)); most = std::max(most, ninput[k]); }
        if (noutput <= 0 && most <= 0) return false;
        if (nout == 0 && most <= 0 && !flush) return false;          // a sink with nothing to read (woken again by its neighbour)
        for (int k = 0; k < nin; k++) {
            // the items are read in place: run_tpb reserved every edge for the whole run, so an append by the producer never
            // moves them; only the counters and the tags of the window are snapshot
            edge* e = rin[k];
            pin[k].item = e->item;
            inp[k] = e->data.data() + e->head;
            pin[k].nread = e->nread; pin[k].nwritten = e->nread + (uint64_t)ninput[k];
            for (auto& t : e->tags) if (t.offset >= e->nread && t.offset < e->nread + (uint64_t)ninput[k]) pin[k].tags.push_back(t);
        }
    }
    for (int k = 0; k < nin; k++) b.mock_in[k] = &pin[k];
    for (int k = 0; k < nout; k++) { pout[k].item = b.output_signature()->sizeof_stream_item(k); b.mock_out[k].assign(1, &pout[k]); }
    gr_vector_const_void_star in(nin);
    for (int k = 0; k < nin; k++) in[k] = inp[k];
    std::vector<std::vector<char>> obuf(nout);
    gr_vector_void_star out(nout);
    for (int k = 0; k < nout; k++) { obuf[k].assign((size_t)std::max(noutput, 1) * pout[k].item, 0); out[k] = obuf[k].data(); }
    b.mock_consumed = 0;
    const size_t tagsBefore = b.mock_tags_added.size(), msgBefore = b.mock_messages.size();
    const auto tw0 = std::chrono::steady_clock::now();
    const int produced = b.general_work(noutput, ninput, in, out);
    const double wms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
    if (produced < 0 || produced > noutput || b.mock_consumed < 0) { std::cerr << b.name() << ": bad accounting" << std::endl; exit(3); }
    const bool moved = b.mock_consumed > 0 || produced > 0 || b.mock_tags_added.size() != tagsBefore || b.mock_messages.size() != msgBefore;
    {
        std::unique_lock<std::mutex> lk(T.mu);
        for (int k = 0; k < nin; k++) {
            edge* e = rin[k];
            if ((size_t)b.mock_consumed > e->avail()) { std::cerr << b.name() << ": consumed more than available" << std::endl; exit(3); }
            e->head += (size_t)b.mock_consumed * e->item;            // (no compaction in this mode: readers hold pointers into the edge)
            e->nread += b.mock_consumed;
            if (e->tags.size() > 64) {
                size_t q = 0;
                for (auto& t : e->tags) if (t.offset >= e->nread) e->tags[q++] = t;
                e->tags.resize(q);
            }
        }
        for (int k = 0; k < nout; k++) {
            const size_t bytes = (size_t)produced * pout[k].item;
            for (edge* e : rout[k]) {
                e->data.insert(e->data.end(), obuf[k].begin(), obuf[k].begin() + (ptrdiff_t)bytes);
                e->nwritten += produced;
                for (auto& t : pout[k].tags) e->tags.push_back(t);
            }
            b.mock_written[k] += produced;
        }
        auto& acc = g_wall[b.name()];
        acc.first += wms; acc.second++;
        if (moved) {
            T.epoch++;
            T.lastNs = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - T.t0).count();
        }
    }
    if (moved) T.cv.notify_all();
    return moved;
}

static double run_tpb(gr::block* const order[5], int maxCall)
{
    Tpb T;
    std::vector<std::vector<edge*>> rin(5);
    std::vector<std::vector<std::vector<edge*>>> rout(5);
    for (int i = 0; i < 5; i++) { rin[i] = order[i]->mock_in; rout[i] = order[i]->mock_out; }
    // room for everything the run can append: the samples once per edge, the soft bits at <= 416 per 72 samples and antenna
    size_t nsamp = 0;
    for (edge* e : rin[0]) nsamp = std::max(nsamp, e->avail());
    for (int i = 0; i < 5; i++)
        for (auto& port : rout[i])
            for (edge* e : port) e->data.reserve(e->data.size() + (size_t)e->item * (e->item == 4 ? nsamp * 12 + (1u << 20) : nsamp * 2 + (1u << 20)));
    T.t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 0; i < 5; i++)
        th.emplace_back([&, i]() {
            int still = 0;                                       // calls in a row that moved nothing
            while (!T.stop.load()) {
                if (call_tpb(*order[i], rin[i], rout[i], maxCall, T)) { still = 0; continue; }
                // a state machine may change state without consuming or producing (RDTAG -> FORMAT ...): GNU Radio's executor
                // calls such a block again at once; only after three such calls does the thread wait for a neighbour
                if (++still < 3) continue;
                // wait for a neighbour to move something: poll the progress counter for a while (a futex wake-up costs tens of
                // microseconds, a frame is worth about a hundred), then sleep on the condition variable
                const uint64_t e0 = T.epoch.load();
                const auto tw = std::chrono::steady_clock::now();
                bool woke = false;
                while (std::chrono::steady_clock::now() - tw < std::chrono::microseconds(300))
                    if (T.epoch.load(std::memory_order_relaxed) != e0 || T.stop.load() || T.flush.load()) { woke = true; break; }
                if (!woke) {
                    std::unique_lock<std::mutex> lk(T.mu);
                    T.cv.wait_for(lk, std::chrono::microseconds(200));
                }
                still = 0;
            }
        });
    uint64_t seen = T.epoch.load();
    int quiet = 0;
    long long idleNs = 0, lastBeforeFlush = 0;
    while (true) {
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
        const uint64_t e = T.epoch.load();
        if (e != seen) { seen = e; quiet = 0; continue; }
        if (++quiet < 25) continue;                                  // 50 ms without a move
        if (!T.flush.load()) {
            // the quiet 50 ms that led here are the scheduler waiting, not the chain working: they do not count
            lastBeforeFlush = T.lastNs.load();
            idleNs = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - T.t0).count() - lastBeforeFlush;
            T.flush = true; quiet = 0; T.cv.notify_all();
            continue;
        }
        break;
    }
    T.stop = true;
    T.cv.notify_all();
    for (auto& t : th) t.join();
    for (int i = 0; i < 5; i++) { order[i]->mock_in = rin[i]; order[i]->mock_out = rout[i]; }
    const long long last = T.lastNs.load();
    return (double)(last > lastBeforeFlush ? last - idleNs : last) * 1e-6;      // (items that only moved in the closing phase)
}

static std::string show(const pmt::pmt_t& v)
{
    char buf[64];
    if (auto p = dynamic_cast<const pmt::p_long*>(v.get())) return std::to_string(p->v);
    if (auto p = dynamic_cast<const pmt::p_real*>(v.get())) { snprintf(buf, sizeof buf, "%.9g", p->v); return buf; }
    if (auto p = dynamic_cast<const pmt::p_c32v*>(v.get())) {
        double s = 0;
        for (auto& c : p->v) s += std::abs(c);
        snprintf(buf, sizeof buf, "c32[%zu]:%.6g", p->v.size(), s);
        return buf;
    }
    return "?";
}

int main(int argc, char** argv)
{
    if (argc != 9) { std::cerr << "usage: run_chain NANT MUPOS MUGID SEED MAXCALL IFDEBUG INDIR OUTFILE" << std::endl; return 2; }
    const int nant = atoi(argv[1]), mupos = atoi(argv[2]), mugid = atoi(argv[3]), maxCall = atoi(argv[5]);
    rng_state = (uint32_t)atoi(argv[4]) * 2654435761u + 1u;
    const bool dbg = atoi(argv[6]) != 0;
    const std::string dir = argv[7];
    using namespace gr::ieee80211;
    std::shared_ptr<gr::block> trig, syn, sig, dem, dec;
    try {
        trig = trigger::make();
        syn = sync::make();
        if (nant == 2) { sig = signal2::make(); dem = demod2::make(); }
        else { sig = signal::make(); dem = demod::make(mupos, mugid); }
        dec = decode::make(dbg);
    } catch (const std::exception& e) {
        std::cerr << "make() failed: " << e.what() << std::endl;
        return 4;
    }
    Graph g;
    const std::vector<char> preac = slurp(dir + "/preac.f32"), preconj = slurp(dir + "/preconj.c64"), s0 = slurp(dir + "/sig.c64");
    g.source(preac, 4, *trig, 0);
    g.connect(*trig, 0, *syn, 0);
    g.source(preconj, 8, *syn, 1);
    g.source(s0, 8, *syn, 2);
    g.connect(*syn, 0, *sig, 0);
    g.source(s0, 8, *sig, 1);
    if (nant == 2) g.source(slurp(dir + "/sig1.c64"), 8, *sig, 2);
    for (int a = 0; a < nant; a++) g.connect(*sig, a, *dem, a);
    g.connect(*dem, 0, *dec, 0);

    gr::block* order[5] = { trig.get(), syn.get(), sig.get(), dem.get(), dec.get() };
    const auto t0 = std::chrono::steady_clock::now();
    const bool tpb = getenv("RUN_CHAIN_TPB") && atoi(getenv("RUN_CHAIN_TPB")) != 0;
    double ms;
    if (tpb) ms = run_tpb(order, maxCall);                        // one thread per block; ms = up to the last item that moved
    else {
        for (int idle = 0; idle < 3;) {
            bool moved = false;
            for (gr::block* b : order)
                for (int r = 1 + (int)(rnd() % 3u); r > 0; r--) moved |= call(*b, maxCall, idle > 0);
            idle = moved ? 0 : idle + 1;
        }
        for (gr::block* b : order) b->stop();                     // (GNU Radio calls stop() on every block: decode publishes what is in flight)
        ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (tpb) for (gr::block* b : order) b->stop();
    std::cout << "run_chain: " << s0.size() / 8 << " samples, " << dec->mock_messages.size() << " messages, " << ms << " ms wall"
              << (tpb ? " (thread per block)" : " (one thread, round robin)") << std::endl;
    for (auto& kv : g_wall) std::cout << "run_chain:   " << kv.first << " " << kv.second.first << " ms in " << kv.second.second << " calls" << std::endl;

    FILE* f = fopen(argv[8], "w");
    if (!f) return 2;
    for (auto& m : dec->mock_messages) {
        const pmt::pmt_t blob = pmt::cdr(m.second);
        const uint8_t* p = (const uint8_t*)pmt::blob_data(blob);
        const size_t n = pmt::blob_length(blob);
        if ((size_t)pmt::to_long(pmt::dict_ref(pmt::car(m.second), pmt::mp("len"), pmt::from_long(-1))) != n) { std::cerr << "len meta" << std::endl; return 3; }
        fprintf(f, "MSG ");
        for (size_t k = 0; k < n; k++) fprintf(f, "%02x", p[k]);
        fprintf(f, "\n");
    }
    for (gr::block* b : order)
        for (auto& t : b->mock_tags_added)
            fprintf(f, "TAG %s %llu %s %s\n", b->name().c_str(), (unsigned long long)t.offset, pmt::symbol_to_string(t.key).c_str(), show(t.value).c_str());
    fclose(f);
    return 0;
}
