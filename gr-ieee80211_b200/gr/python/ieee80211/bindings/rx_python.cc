/* rx_python.cc -- pybind11 binding of gr::ieee80211::rx (gr/include/gnuradio/ieee80211/rx.h), to sit next to the
 * reference's python/ieee80211/bindings/decode_python.cc: add `void bind_rx(py::module&);` and `bind_rx(m);` to
 * python_bindings.cc and this file to the bindings' CMakeLists.txt.  Gives Python / GRC
 *     from gnuradio import ieee80211;  blk = ieee80211.rx(nant=1, mupos=0, mugid=2, ifdebug=False)
 * (grc/ieee80211_rx.block.yml).  The seven reference blocks keep their own, unchanged bindings.
 */
#include <pybind11/pybind11.h>

#include <memory>

namespace py = pybind11;

#include <gnuradio/ieee80211/rx.h>

void bind_rx(py::module& m)
{
    using rx = ::gr::ieee80211::rx;

    py::class_<rx, gr::block, gr::basic_block, std::shared_ptr<rx>>(
        m, "rx", "802.11 OFDM receive chain (presiso ... decode) as one sink block on the GPU; PDUs on message port 'out'")
        .def(py::init(&rx::make), py::arg("nant") = 1, py::arg("mupos") = 0, py::arg("mugid") = 2, py::arg("ifdebug") = false,
             "nant: 1 or 2 receive antennas; mupos / mugid as demod(mupos, mugid); ifdebug as decode(ifdebug)");
}
