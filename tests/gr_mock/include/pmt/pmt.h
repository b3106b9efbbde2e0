// TEST INFRASTRUCTURE: a miniature of GNU Radio's pmt -- just the calls the block shells (gr-ieee80211_b200/gr/lib) make,
// with the real library's names and signatures, so the shells compile and run without GNU Radio installed.
#pragma once
#include <complex>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pmt {

struct pmt_base { virtual ~pmt_base() {} };
typedef std::shared_ptr<pmt_base> pmt_t;

struct p_nil : pmt_base {};
struct p_bool : pmt_base { bool v; explicit p_bool(bool b) : v(b) {} };
struct p_sym : pmt_base { std::string s; explicit p_sym(const std::string& x) : s(x) {} };
struct p_long : pmt_base { long v; explicit p_long(long x) : v(x) {} };
struct p_real : pmt_base { double v; explicit p_real(double x) : v(x) {} };
struct p_c32v : pmt_base { std::vector<std::complex<float>> v; };
struct p_blob : pmt_base { std::vector<uint8_t> v; };
struct p_pair : pmt_base { pmt_t a, b; };
struct p_dict : pmt_base { std::vector<std::pair<pmt_t, pmt_t>> kv; };
struct p_list : pmt_base { std::vector<pmt_t> v; };

inline const pmt_t PMT_NIL = std::make_shared<p_nil>();
inline const pmt_t PMT_F = std::make_shared<p_bool>(false);
inline const pmt_t PMT_T = std::make_shared<p_bool>(true);

template <class T> const T& as(const pmt_t& p, const char* what)
{
    const T* q = dynamic_cast<const T*>(p.get());
    if (!q) throw std::runtime_error(std::string("pmt: wrong type, wanted ") + what);
    return *q;
}

inline pmt_t mp(const std::string& s) { return std::make_shared<p_sym>(s); }
inline pmt_t mp(const char* s) { return std::make_shared<p_sym>(s); }
inline pmt_t intern(const std::string& s) { return mp(s); }
inline std::string symbol_to_string(const pmt_t& p) { return as<p_sym>(p, "symbol").s; }
inline pmt_t from_long(long v) { return std::make_shared<p_long>(v); }
inline long to_long(const pmt_t& p) { return as<p_long>(p, "long").v; }
inline pmt_t from_float(double v) { return std::make_shared<p_real>(v); }
inline pmt_t from_double(double v) { return std::make_shared<p_real>(v); }
inline float to_float(const pmt_t& p) { return (float)as<p_real>(p, "real").v; }
inline double to_double(const pmt_t& p) { return as<p_real>(p, "real").v; }
inline pmt_t init_c32vector(size_t n, const std::vector<std::complex<float>>& d)
{
    auto p = std::make_shared<p_c32v>();
    p->v.assign(d.begin(), d.begin() + n);
    return p;
}
inline const std::vector<std::complex<float>> c32vector_elements(const pmt_t& p) { return as<p_c32v>(p, "c32vector").v; }
inline pmt_t make_blob(const void* buf, size_t n)
{
    auto p = std::make_shared<p_blob>();
    p->v.assign((const uint8_t*)buf, (const uint8_t*)buf + n);
    return p;
}
inline const void* blob_data(const pmt_t& p) { return as<p_blob>(p, "blob").v.data(); }
inline size_t blob_length(const pmt_t& p) { return as<p_blob>(p, "blob").v.size(); }
inline pmt_t cons(const pmt_t& a, const pmt_t& b) { auto p = std::make_shared<p_pair>(); p->a = a; p->b = b; return p; }
inline pmt_t car(const pmt_t& p) { return as<p_pair>(p, "pair").a; }
inline pmt_t cdr(const pmt_t& p) { return as<p_pair>(p, "pair").b; }
inline pmt_t make_dict() { return std::make_shared<p_dict>(); }
inline pmt_t dict_add(const pmt_t& d, const pmt_t& k, const pmt_t& v)
{
    auto p = std::make_shared<p_dict>(as<p_dict>(d, "dict"));
    for (auto& kv : p->kv) if (symbol_to_string(kv.first) == symbol_to_string(k)) { kv.second = v; return p; }
    p->kv.emplace_back(k, v);
    return p;
}
inline pmt_t dict_ref(const pmt_t& d, const pmt_t& k, const pmt_t& dflt)
{
    for (auto& kv : as<p_dict>(d, "dict").kv) if (symbol_to_string(kv.first) == symbol_to_string(k)) return kv.second;
    return dflt;
}
// the real dict is an association list that dict_add conses onto: dict_items lists the newest key first
inline pmt_t dict_items(const pmt_t& d)
{
    auto l = std::make_shared<p_list>();
    const auto& kv = as<p_dict>(d, "dict").kv;
    for (auto it = kv.rbegin(); it != kv.rend(); ++it) l->v.push_back(cons(it->first, it->second));
    return l;
}
inline size_t length(const pmt_t& p)
{
    if (auto l = dynamic_cast<const p_list*>(p.get())) return l->v.size();
    if (auto d = dynamic_cast<const p_dict*>(p.get())) return d->kv.size();
    if (auto c = dynamic_cast<const p_c32v*>(p.get())) return c->v.size();
    if (auto b = dynamic_cast<const p_blob*>(p.get())) return b->v.size();
    throw std::runtime_error("pmt: length of a non-sequence");
}
inline pmt_t nth(size_t n, const pmt_t& list) { return as<p_list>(list, "list").v.at(n); }
inline bool is_null(const pmt_t& p) { return dynamic_cast<const p_nil*>(p.get()) != nullptr; }

}  // namespace pmt
