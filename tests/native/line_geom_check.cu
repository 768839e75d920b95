// Host-side check of the line table's key arithmetic (kmers.anno_b200/csrc/ka_line.cuh), compiled by
// tests/test_host.py with nvcc and run on the CPU: the mixer is a bijection, the c-way split of the top bits is a
// bijection, and (line, tag) identifies the key — what quotienting relies on.  Prints "ok" or the first violation.
#include <cstdio>
#include <cstdint>
#include <vector>
#include "ka_line.cuh"

using namespace ka;

static int fail(const char* what, unsigned long long a, unsigned long long b) {
    printf("FAIL %s %llu %llu\n", what, a, b);
    return 1;
}

int main() {
    // 1. the Feistel mixer permutes (bh, bl)-bit pairs: exhaustive for small widths
    for (uint32_t bh = 3; bh <= 9; bh += 3)
        for (uint32_t bl = 3; bl <= 10; bl += 7) {
            std::vector<uint8_t> seen((size_t)1 << (bh + bl), 0);
            for (uint32_t H = 0; H < (1u << bh); H++)
                for (uint32_t Lo = 0; Lo < (1u << bl); Lo++) {
                    uint32_t L = H, R = Lo;
                    line_mix(L, R, bh, bl);
                    if (L >> bh || R >> bl) return fail("mix range", H, Lo);
                    size_t i = ((size_t)L << bl) | R;
                    if (seen[i]) return fail("mix collision", H, Lo);
                    seen[i] = 1;
                }
        }
    // 2. U -> (hi, idx) is a bijection for every c of the layout (u = 12), and for the power-of-two case (u = 3, c = 8)
    for (uint32_t c = 8; c < 16; c++) {
        const uint32_t u = 12, inv_c = (65536u + c - 1) / c;
        std::vector<uint8_t> seen((size_t)c << (u - 3), 0);
        for (uint32_t U = 0; U < (1u << u); U++) {
            const uint32_t x = U * c, hi = x >> u, idx = ((x & ((1u << u) - 1u)) * inv_c) >> 16;
            if (hi >= c || idx >= (1u << (u - 3))) return fail("split range", c, U);
            if (idx != (x & ((1u << u) - 1u)) / c) return fail("split division", c, U);
            size_t i = ((size_t)hi << (u - 3)) | idx;
            if (seen[i]) return fail("split collision", c, U);
            seen[i] = 1;
        }
    }
    // 3. (line, tag) identifies the key: exhaustive over a whole small key space (4 symbols, K = 8 and K = 7)
    for (int K = 7; K <= 8; K++) {
        LineTable t{};
        const uint32_t nsym = 4, Kh = K / 2, Kl = K - Kh;
        t.bh = 2 * Kh; t.bl = 2 * Kl;                     // 4^k values = 2k bits
        t.u = 3; t.c = 8; t.inv_c = (65536u + 7) / 8;
        t.la = t.bh - t.u;
        t.s = t.la + 2;                                   // a = 2 bits of R in the line index
        t.a = t.s - t.la; t.r = t.bl - t.a;
        t.n_lines = t.c << t.s;
        t.radix = nsym; t.Kh = Kh; t.Kl = Kl; t.K = K;
        if ((t.u - 3) + t.r > TAG_REM_BITS) return fail("test geometry", t.r, 0);
        std::vector<uint8_t> seen(((size_t)t.n_lines * 4) << TAG_REM_BITS, 0);
        for (uint32_t H = 0; H < (1u << t.bh); H++)
            for (uint32_t Lo = 0; Lo < (1u << t.bl); Lo++) {
                uint32_t sector, tag;
                line_locate(t, H, Lo, sector, tag);
                if (sector >= t.n_lines * 4u) return fail("sector range", H, Lo);
                if (!(tag & TAG_VALID) || (tag & TAG_FLAG) || tag >> 16) return fail("tag bits", H, Lo);
                // the low two bits of the remainder are the home sector: line and remainder together name the key
                size_t i = (((size_t)(sector >> 2)) << TAG_REM_BITS) | (tag & ((1u << TAG_REM_BITS) - 1u));
                if (((tag & 3u) != (sector & 3u))) return fail("home sector bits", H, Lo);
                if (seen[i]) return fail("locate collision", H, Lo);
                seen[i] = 1;
            }
    }
    // 4. filter word index stays in range and the two bits are inside the word
    for (uint32_t n_filt : {1u, 7u, 4096u, 18874368u})
        for (uint32_t k = 0; k < 100000; k++) {
            const uint32_t fh = line_filter_hash(k * 2654435761u, k ^ 0x5bd1e995u);
            if (line_filter_word(fh, n_filt) >= n_filt) return fail("filter word", n_filt, k);
            if (line_filter_bits(fh) == 0) return fail("filter bits", n_filt, k);
        }
    printf("ok\n");
    return 0;
}
