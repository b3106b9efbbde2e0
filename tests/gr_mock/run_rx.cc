// TEST INFRASTRUCTURE: plays the GNU Radio scheduler for the compiled sink block gr::ieee80211::rx
// (gr-ieee80211_b200/gr/lib/rx_impl.cc) over the miniature runtime of tests/gr_mock/include: feeds the capture in
// pseudo-random pieces, calls stop() at the end, dumps the published messages.
//   run_rx NANT MUPOS MUGID SEED MAXCALL IFDEBUG INDIR OUTFILE      (INDIR holds sig.c64 [, sig1.c64])
#include <gnuradio/ieee80211/rx.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

static std::vector<char> slurp(const std::string& p)
{
    std::ifstream f(p, std::ios::binary);
    if (!f) { std::cerr << "cannot read " << p << std::endl; exit(2); }
    return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv)
{
    if (argc != 9) { std::cerr << "usage: run_rx NANT MUPOS MUGID SEED MAXCALL IFDEBUG INDIR OUTFILE" << std::endl; return 2; }
    const int nant = atoi(argv[1]), maxCall = atoi(argv[5]);
    uint32_t rng = (uint32_t)atoi(argv[4]) * 2654435761u + 1u;
    const std::string dir = argv[7];
    std::shared_ptr<gr::block> blk;
    try {
        blk = gr::ieee80211::rx::make(nant, atoi(argv[2]), atoi(argv[3]), atoi(argv[6]) != 0);
    } catch (const std::exception& e) {
        std::cerr << "make() failed: " << e.what() << std::endl;
        return 4;
    }
    std::vector<std::vector<char>> sig;
    sig.push_back(slurp(dir + "/sig.c64"));
    if (nant == 2) sig.push_back(slurp(dir + "/sig1.c64"));
    std::vector<gr::mock::edge> edges(nant);
    for (int a = 0; a < nant; a++) { edges[a].item = 8; blk->mock_in[a] = &edges[a]; }
    const size_t total = sig[0].size() / 8;
    const auto t0 = std::chrono::steady_clock::now();
    blk->start();
    for (size_t done = 0; done < total;) {
        rng = rng * 1664525u + 1013904223u;
        const size_t n = std::min<size_t>(total - done, 1 + (rng >> 8) % (uint32_t)maxCall);
        gr_vector_int ninput(nant, (int)n);
        gr_vector_const_void_star in(nant);
        gr_vector_void_star out;
        for (int a = 0; a < nant; a++) in[a] = sig[a].data() + done * 8;
        blk->mock_consumed = 0;
        const int produced = blk->general_work((int)n, ninput, in, out);
        if (produced != 0 || blk->mock_consumed != (int)n) { std::cerr << "rx: bad accounting" << std::endl; return 3; }
        for (int a = 0; a < nant; a++) edges[a].nread += n;
        done += n;
    }
    blk->stop();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "run_rx: " << total << " samples, " << blk->mock_messages.size() << " messages, " << ms << " ms wall" << std::endl;
    FILE* f = fopen(argv[8], "w");
    if (!f) return 2;
    for (auto& m : blk->mock_messages) {
        const pmt::pmt_t blob = pmt::cdr(m.second);
        const uint8_t* p = (const uint8_t*)pmt::blob_data(blob);
        const size_t n = pmt::blob_length(blob);
        if ((size_t)pmt::to_long(pmt::dict_ref(pmt::car(m.second), pmt::mp("len"), pmt::from_long(-1))) != n) { std::cerr << "len meta" << std::endl; return 3; }
        fprintf(f, "MSG ");
        for (size_t k = 0; k < n; k++) fprintf(f, "%02x", p[k]);
        fprintf(f, "\n");
    }
    fclose(f);
    return 0;
}
