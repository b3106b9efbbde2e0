// TEST INFRASTRUCTURE (see pmt/pmt.h): gr::fft::fft_complex_fwd with the members the reference's blocks call
// (lib/signal_impl.cc:38-45,121-124, lib/demod_impl.cc:40,541-547, lib/demod2_impl.cc:38,790-796): forward, unnormalised,
// natural order.  GNU Radio's class runs FFTW3f (absent from the image, version unpinned by the reference); this stand-in
// evaluates the same transform as oracle/oracle_rx.cc's fft64 does -- radix-2 DIT in double, rounded to float once -- so the
// reference blocks and the restated oracle see bit-identical spectra and every later difference is the blocks' own.
#pragma once
#include <gnuradio/gr_complex.h>

#include <cmath>
#include <stdexcept>
#include <vector>

namespace gr {
namespace fft {

class fft_complex_fwd
{
public:
    explicit fft_complex_fwd(int fft_size, int nthreads = 1) : d_n(fft_size), d_in(fft_size), d_out(fft_size), d_re(fft_size), d_im(fft_size)
    {
        (void)nthreads;
        if (fft_size < 2 || (fft_size & (fft_size - 1))) throw std::invalid_argument("mock fft: power-of-two sizes only");
        d_bits = 0;
        while ((1 << d_bits) < d_n) d_bits++;
        d_c.resize(d_n / 2); d_s.resize(d_n / 2); d_rev.resize(d_n);
        for (int i = 0; i < d_n / 2; i++) { d_c[i] = cos(-2.0 * M_PI * i / (double)d_n); d_s[i] = sin(-2.0 * M_PI * i / (double)d_n); }
        for (int i = 0; i < d_n; i++) { int r = 0; for (int b = 0; b < d_bits; b++) r |= ((i >> b) & 1) << (d_bits - 1 - b); d_rev[i] = r; }
    }
    gr_complex* get_inbuf() { return d_in.data(); }
    gr_complex* get_outbuf() { return d_out.data(); }
    int inbuf_length() const { return d_n; }
    int outbuf_length() const { return d_n; }
    void execute()
    {
        double* re = d_re.data(); double* im = d_im.data();
        for (int i = 0; i < d_n; i++) { re[d_rev[i]] = d_in[i].real(); im[d_rev[i]] = d_in[i].imag(); }
        for (int len = 2; len <= d_n; len <<= 1) {
            const int half = len >> 1, step = d_n / len;
            for (int b = 0; b < d_n; b += len)
                for (int k = 0; k < half; k++) {
                    const double wr = d_c[k * step], wi = d_s[k * step];
                    const double xr = re[b + k + half] * wr - im[b + k + half] * wi;
                    const double xi = re[b + k + half] * wi + im[b + k + half] * wr;
                    re[b + k + half] = re[b + k] - xr; im[b + k + half] = im[b + k] - xi;
                    re[b + k] += xr; im[b + k] += xi;
                }
        }
        for (int i = 0; i < d_n; i++) d_out[i] = gr_complex((float)re[i], (float)im[i]);
    }
private:
    int d_n, d_bits;
    std::vector<gr_complex> d_in, d_out;
    std::vector<double> d_re, d_im, d_c, d_s;
    std::vector<int> d_rev;
};

}  // namespace fft
}  // namespace gr
