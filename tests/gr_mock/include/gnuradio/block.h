// TEST INFRASTRUCTURE (see pmt/pmt.h): gr::basic_block / gr::block with the member names, signatures and ACCESS LEVELS of
// GNU Radio 3.10 (add_item_tag / get_tags_in_range are protected there), over a tiny runtime: an edge is a byte queue with
// read / write counters and tags at absolute offsets; tests/gr_mock/run_chain.cc plays the scheduler.
#pragma once
#include <gnuradio/attributes.h>
#include <gnuradio/io_signature.h>
#include <pmt/pmt.h>

// the real <gnuradio/block.h> pulls these in transitively and the reference's blocks rely on it (std::cout, memset,
// std::max_element, std::chrono without their own #include)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

typedef std::vector<int> gr_vector_int;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

namespace gnuradio {
template <class T, class... Args>
std::shared_ptr<T> make_block_sptr(Args&&... args) { return std::shared_ptr<T>(new T(std::forward<Args>(args)...)); }
}  // namespace gnuradio

namespace gr {

struct tag_t {
    uint64_t offset = 0;
    pmt::pmt_t key, value, srcid;
};

namespace mock {
struct edge {
    int item = 1;
    std::vector<char> data;          // items written and not yet consumed (from byte `head` on)
    size_t head = 0;                 // schedulers that do not erase consumed items advance this instead
    uint64_t nread = 0, nwritten = 0;
    std::vector<tag_t> tags;
    size_t avail() const { return (data.size() - head) / (size_t)item; }
};
}  // namespace mock

class basic_block
{
public:
    virtual ~basic_block() {}
    std::string name() const { return d_name; }
    pmt::pmt_t alias_pmt() const { return pmt::mp(d_name); }
    void message_port_register_out(pmt::pmt_t port_id) { d_msg_ports.push_back(pmt::symbol_to_string(port_id)); }
    void message_port_pub(pmt::pmt_t port_id, pmt::pmt_t msg) { mock_messages.emplace_back(pmt::symbol_to_string(port_id), msg); }
    // mock runtime
    std::vector<std::pair<std::string, pmt::pmt_t>> mock_messages;
protected:
    basic_block() {}
    explicit basic_block(const std::string& n) : d_name(n) {}
    std::string d_name;
    std::vector<std::string> d_msg_ports;
};

class block : public basic_block
{
public:
    enum tag_propagation_policy_t { TPP_DONT = 0, TPP_ALL_TO_ALL = 1, TPP_ONE_TO_ONE = 2, TPP_CUSTOM = 3 };
    typedef std::shared_ptr<block> sptr;

    virtual void forecast(int noutput_items, gr_vector_int& ninput_items_required)
    {
        for (auto& r : ninput_items_required) r = noutput_items;
    }
    virtual int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                             gr_vector_void_star& output_items) = 0;
    virtual bool start() { return true; }
    virtual bool stop() { return true; }
    void set_output_multiple(int multiple) { mock_output_multiple = multiple; }
    void consume_each(int how_many_items) { mock_consumed = how_many_items; }
    uint64_t nitems_read(unsigned int which_input) { return mock_in.at(which_input)->nread; }
    uint64_t nitems_written(unsigned int which_output) { return mock_written.at(which_output); }
    tag_propagation_policy_t tag_propagation_policy() { return d_tpp; }
    void set_tag_propagation_policy(tag_propagation_policy_t p) { d_tpp = p; }
    io_signature::sptr input_signature() const { return d_in; }
    io_signature::sptr output_signature() const { return d_out; }

    // mock runtime: input edges (one per port), output edges (fan-out per port), counters
    std::vector<mock::edge*> mock_in;
    std::vector<std::vector<mock::edge*>> mock_out;
    std::vector<uint64_t> mock_written;
    int mock_consumed = 0, mock_output_multiple = 1;

protected:
    block() {}
    block(const std::string& name, io_signature::sptr input_signature, io_signature::sptr output_signature)
        : basic_block(name), d_in(input_signature), d_out(output_signature)
    {
        mock_in.assign((size_t)d_in->max_streams(), nullptr);
        mock_out.assign((size_t)d_out->max_streams(), {});
        mock_written.assign((size_t)d_out->max_streams(), 0);
    }
    void add_item_tag(unsigned int which_output, uint64_t abs_offset, const pmt::pmt_t& key, const pmt::pmt_t& value,
                      const pmt::pmt_t& srcid = pmt::PMT_F)
    {
        tag_t t;
        t.offset = abs_offset; t.key = key; t.value = value; t.srcid = srcid;
        for (mock::edge* e : mock_out.at(which_output)) e->tags.push_back(t);
        mock_tags_added.push_back(t);
    }
    void get_tags_in_range(std::vector<tag_t>& v, unsigned int which_input, uint64_t abs_start, uint64_t abs_end)
    {
        v.clear();
        for (const tag_t& t : mock_in.at(which_input)->tags)
            if (t.offset >= abs_start && t.offset < abs_end) v.push_back(t);
    }
public:
    std::vector<tag_t> mock_tags_added;   // everything this block ever attached (for the test's dump)
private:
    io_signature::sptr d_in, d_out;
    tag_propagation_policy_t d_tpp = TPP_ALL_TO_ALL;
};

}  // namespace gr
