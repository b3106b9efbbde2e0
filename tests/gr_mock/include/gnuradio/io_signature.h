// TEST INFRASTRUCTURE (see pmt/pmt.h): gr::io_signature::make / makev
#pragma once
#include <gnuradio/gr_complex.h>
#include <memory>
#include <vector>

namespace gr {
class io_signature
{
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int min_streams, int max_streams, int sizeof_stream_item)
    {
        return makev(min_streams, max_streams, std::vector<int>(max_streams > 0 ? max_streams : 0, sizeof_stream_item));
    }
    static sptr makev(int min_streams, int max_streams, const std::vector<int>& sizeof_stream_items)
    {
        sptr s(new io_signature);
        s->d_min = min_streams; s->d_max = max_streams; s->d_sizes = sizeof_stream_items;
        return s;
    }
    int min_streams() const { return d_min; }
    int max_streams() const { return d_max; }
    int sizeof_stream_item(int k) const { return d_sizes.empty() ? 0 : d_sizes[k < (int)d_sizes.size() ? k : (int)d_sizes.size() - 1]; }
private:
    int d_min = 0, d_max = 0;
    std::vector<int> d_sizes;
};
}  // namespace gr
