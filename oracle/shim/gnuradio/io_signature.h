// TEST INFRASTRUCTURE (oracle/): stand-in for <gnuradio/io_signature.h> so that the
// reference's lib/cloud80211phy.{h,cc} compile unmodified without GNU Radio.
// The only thing that translation unit needs from GNU Radio is the gr_complex typedef.
#pragma once
#include <algorithm>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <vector>
typedef std::complex<float> gr_complex;
