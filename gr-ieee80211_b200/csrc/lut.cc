// lut.cc -- builds the LUT blob by formula (host).  Nothing here is transcribed from the
// reference's tables; tests compare the result with them (tests/test_lut.py).
#include "lut.h"

#include <math.h>
#include <string.h>

void c8b_lut_build(c8b_lut* L)
{
    memset(L, 0, sizeof(*L));
    L->magic = C8B_LUT_MAGIC;
    L->version = C8B_LUT_VERSION;
    L->bytes = (uint32_t)sizeof(c8b_lut);

    // IEEE 802.11-2016 Eq. (17-8): L-LTF on subcarriers -26..26
    static const int8_t ltf[53] = { 1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 0,
                                    1, -1, -1, 1, 1, -1, 1, -1, 1, -1, -1, -1, -1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, 1, 1 };
    for (int k = -26; k <= 26; k++) {
        L->ltfL[(k + 64) & 63] = (float)ltf[k + 26];
        L->ltfNL[(k + 64) & 63] = (float)ltf[k + 26];
    }
    // Eq. (19-23): HT-LTF = {1,1, L-LTF, -1,-1} on -28..28
    L->ltfNL[64 - 28] = 1.f; L->ltfNL[64 - 27] = 1.f; L->ltfNL[27] = -1.f; L->ltfNL[28] = -1.f;
    // VHT 2-LTF frame, second LTF through P-matrix row 2: data tones x(-1), pilot tones (R matrix) x(+1);
    // relative to ltfNL*(-1)... the table a receiver divides by keeps the data sign and flips the pilots.
    for (int i = 0; i < 64; i++) L->ltfNL22[i] = L->ltfNL[i];
    const int pil[4] = { 7, 21, 43, 57 };
    for (int q = 0; q < 4; q++) L->ltfNL22[pil[q]] = -L->ltfNL[pil[q]];

    // pilot polarity: scrambler x^7+x^4+1 from all ones; 0 -> +1, 1 -> -1 (17.3.5.10)
    int st = 0x7f;
    for (int i = 0; i < 127; i++) {
        int fb = ((st >> 6) ^ (st >> 3)) & 1;
        st = ((st << 1) & 0x7e) | fb;
        L->pilotP[i] = fb ? -1.f : 1.f;
    }
    for (int k = 0; k < 64; k++) {
        L->twr[k] = (float)cos(-2.0 * M_PI * k / 64.0);
        L->twi[k] = (float)sin(-2.0 * M_PI * k / 64.0);
        if (k < 32) { L->twdr[k] = cos(-2.0 * M_PI * k / 64.0); L->twdi[k] = sin(-2.0 * M_PI * k / 64.0); }
    }
    // legacy interleaver 17.3.5.7: k -> i -> j; deinterleaver scatter map[j] = k
    const int nbL[4] = { 1, 2, 4, 6 };
    for (int m = 0; m < 4; m++) {
        int ncbps = 48 * nbL[m], s = nbL[m] / 2 > 1 ? nbL[m] / 2 : 1;
        for (int k = 0; k < ncbps; k++) {
            int i = (ncbps / 16) * (k % 16) + k / 16;
            int j = s * (i / s) + (i + ncbps - (16 * i) / ncbps) % s;
            L->deintL[m][j] = (uint16_t)k;
        }
    }
    // HT/VHT 20 MHz interleaver 19.3.11.8.3: N_COL 13, N_ROW 4 N_BPSCS, N_ROT 11
    const int nbN[5] = { 1, 2, 4, 6, 8 };
    for (int iss = 1; iss <= 2; iss++)
        for (int m = 0; m < 5; m++) {
            int n = 52 * nbN[m], s = nbN[m] / 2 > 1 ? nbN[m] / 2 : 1, nrow = 4 * nbN[m];
            int rot = (((iss - 1) * 2) % 3 + 3 * ((iss - 1) / 3)) * 11 * nbN[m];
            for (int k = 0; k < n; k++) {
                int i = nrow * (k % 13) + k / 13;
                int j = s * (i / s) + (i + n - (13 * i) / n) % s;
                int r = ((j - rot) % n + n) % n;
                L->deintNL[iss - 1][m][r] = (uint16_t)k;
            }
        }
    // data-tone numbering: bins in -26..26 (-28..28) order without DC and pilots
    for (int i = 0; i < 64; i++) { L->sigDemap[i] = -1; L->binToDataL[i] = 255; L->binToDataNL[i] = 255; }
    int d = 0;
    for (int k = -26; k <= 26; k++) {
        if (k == 0 || k == -21 || k == -7 || k == 7 || k == 21) continue;
        L->binToDataL[(k + 64) & 63] = (uint8_t)d;
        L->sigDemap[(k + 64) & 63] = (int8_t)L->deintL[0][d];
        d++;
    }
    d = 0;
    for (int k = -28; k <= 28; k++) {
        if (k == 0 || k == -21 || k == -7 || k == 7 || k == 21) continue;
        L->binToDataNL[(k + 64) & 63] = (uint8_t)d++;
    }
    // K=7 code g0=133o g1=171o.  Register: newest input at bit 0, state bit 5 (previous input) at
    // bit 1 ... state bit 0 (oldest) at bit 6.  Output class of the transition 2k --input 0--> k.
    for (int k = 0; k < 32; k++) {
        int s6 = 2 * k, reg = 0;
        for (int q = 0; q < 6; q++) reg |= ((s6 >> (5 - q)) & 1) << (q + 1);
        int o0 = __builtin_popcount(reg & 0155) & 1, o1 = __builtin_popcount(reg & 0117) & 1;
        L->bmClass[k] = (uint8_t)(o0 * 2 + o1);
    }
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        L->crc32tab[i] = c;
    }
    // zero-byte advance of the (reflected) CRC register is linear: columns of Z^(64*2^p)
    for (int i = 0; i < 32; i++) {
        uint32_t c = 1u << i;
        for (int p = 0; p < 6; p++) {
            const int nb = p == 0 ? 64 : 64 << (p - 1);          // advance from Z^(64*2^(p-1)) to Z^(64*2^p)
            for (int k = 0; k < nb; k++) c = L->crc32tab[c & 0xff] ^ (c >> 8);
            L->crcZ[p][i] = c;
        }
    }
    L->pair01[0] = 0.0f;
    L->pair01[1] = 1.0f;
    // per-thread demap entries: read the closed form (base + rotation) off the maps built above
    for (int mode = 0; mode < 9; mode++) {
        const bool leg = mode < 4;
        const int nb = leg ? nbL[mode] : nbN[mode - 4], s = nb / 2 > 1 ? nb / 2 : 1, ncol = leg ? 16 : 13;
        const uint16_t* map = leg ? L->deintL[mode] : L->deintNL[0][mode - 4];
        for (int j = 0; j < 8; j++)
            for (int k2 = 0; k2 < 8; k2++) {
                const int dd = leg ? L->binToDataL[j + 8 * k2] : L->binToDataNL[j + 8 * k2];
                uint16_t e = 0xFFFF;
                if (dd != 255) {
                    int base = map[dd * nb];
                    for (int c = 1; c < s; c++) if (map[dd * nb + c] < base) base = map[dd * nb + c];
                    e = (uint16_t)(base | (((map[dd * nb] - base) / ncol) << 9));
                }
                L->demapTab[mode][8 * j + k2] = e;
            }
    }
    for (int m = 0; m < 5; m++)
        for (int a = 0; a < 2; a++) {
            const int nb = nbN[m], s = nb / 2 > 1 ? nb / 2 : 1;
            const uint16_t* map = L->deintNL[a][m];
            for (int j = 0; j < 8; j++)
                for (int k2 = 0; k2 < 8; k2++) {
                    const int dd = L->binToDataNL[j + 8 * k2];
                    uint16_t e = 0xFFFF;
                    if (dd != 255) {
                        int base = map[dd * nb];
                        for (int c = 1; c < s; c++) if (map[dd * nb + c] < base) base = map[dd * nb + c];
                        const int R = (map[dd * nb] - base) / 13, p0 = base + s * (base / s) + a * s;
                        e = (uint16_t)(p0 | (R << 10) | ((base % s) << 12));
                    }
                    L->demapTab2[m][a][8 * j + k2] = e;
                }
        }
    for (int j = 0; j < 8; j++)
        for (int k1 = 0; k1 < 8; k1++) {
            L->tw8[k1 >> 1][j][2 * (k1 & 1)] = L->twr[(j * k1) & 63];
            L->tw8[k1 >> 1][j][2 * (k1 & 1) + 1] = L->twi[(j * k1) & 63];
        }
}
