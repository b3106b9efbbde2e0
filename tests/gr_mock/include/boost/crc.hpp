// TEST INFRASTRUCTURE (see pmt/pmt.h): boost::crc_32_type with the three members the reference's decode block calls
// (lib/decode_impl.cc:372-374,451-453): the standard reflected CRC-32 (poly 0x04C11DB7, init / final xor 0xFFFFFFFF),
// which is what boost::crc_optimal<32, 0x04C11DB7, 0xFFFFFFFF, 0xFFFFFFFF, true, true> computes.
#pragma once
#include <cstddef>
#include <cstdint>

namespace boost {
class crc_32_type
{
public:
    typedef uint32_t value_type;
    crc_32_type() { reset(); }
    void reset() { d_rem = 0xFFFFFFFFu; }
    void process_byte(unsigned char b)
    {
        d_rem ^= b;
        for (int k = 0; k < 8; k++) d_rem = (d_rem & 1u) ? (0xEDB88320u ^ (d_rem >> 1)) : (d_rem >> 1);
    }
    void process_bytes(const void* buffer, std::size_t n)
    {
        const unsigned char* p = static_cast<const unsigned char*>(buffer);
        const uint32_t* t = table();
        uint32_t r = d_rem;
        for (std::size_t i = 0; i < n; i++) r = t[(r ^ p[i]) & 0xFFu] ^ (r >> 8);
        d_rem = r;
    }
    void process_block(const void* b, const void* e) { process_bytes(b, (std::size_t)(static_cast<const char*>(e) - static_cast<const char*>(b))); }
    value_type checksum() const { return d_rem ^ 0xFFFFFFFFu; }
private:
    static const uint32_t* table()
    {
        static uint32_t t[256];
        static bool init = [] {
            for (uint32_t i = 0; i < 256; i++) {
                uint32_t c = i;
                for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
                t[i] = c;
            }
            return true;
        }();
        (void)init;
        return t;
    }
    uint32_t d_rem;
};
}  // namespace boost
