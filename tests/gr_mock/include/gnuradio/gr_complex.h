// TEST INFRASTRUCTURE (see pmt/pmt.h): gr_complex, as <gnuradio/gr_complex.h> of GNU Radio 3.10 defines it
#pragma once
#include <complex>
typedef std::complex<float> gr_complex;
typedef std::complex<double> gr_complexd;
