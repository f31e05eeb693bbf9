// tools/csv_exhaustive.cpp -- development aid: checks smart_csv_format_f32 (include/smart_b200_io.h)
// against the C library's exact "%.6e" for EVERY float32 bit pattern, and smart_csv_parse_f32 against
// strtod on every text it produced.  ~10 minutes on 8 cores.
//   g++ -O2 -std=c++17 -I include -o build_exp/csv_exhaustive tools/csv_exhaustive.cpp -L smartpy_b200 -lsmart_b200 -lpthread
//   LD_LIBRARY_PATH=smartpy_b200 build_exp/csv_exhaustive
#include "smart_b200_io.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>

int main(int argc, char **argv)
{
    const int workers = argc > 1 ? atoi(argv[1]) : static_cast<int>(std::thread::hardware_concurrency());
    const uint64_t total = 1ull << 32, block = 1ull << 20;
    std::atomic<uint64_t> next{0}, bad_format{0}, bad_parse{0}, checked{0};
    auto work = [&]() {
        std::vector<float> v(block);
        std::vector<char> text(static_cast<size_t>(smart_csv_bound(block, 1)));
        std::vector<float> back(block);
        const int32_t wanted = 0;
        char ref[32];
        for (;;) {
            const uint64_t first = next.fetch_add(block);
            if (first >= total) break;
            for (uint64_t i = 0; i < block; ++i) {
                const uint32_t bits = static_cast<uint32_t>(first + i);
                memcpy(&v[i], &bits, 4);
            }
            const int64_t n = smart_csv_format_f32(v.data(), block, 1, 1, text.data(), static_cast<int64_t>(text.size()), 1);
            const char *p = text.data();
            for (uint64_t i = 0; i < block; ++i) {
                const char *nl = static_cast<const char *>(memchr(p, '\n', text.data() + n - p));
                const float f = v[i];
                int len;
                if (f != f) len = snprintf(ref, sizeof ref, "nan");
                else if (isinf(f)) len = snprintf(ref, sizeof ref, f < 0 ? "-inf" : "inf");
                else len = snprintf(ref, sizeof ref, "%.6e", static_cast<double>(f));
                if (nl - p != len || memcmp(p, ref, len) != 0) {
                    if (bad_format.fetch_add(1) < 10) fprintf(stderr, "format %08x: %.*s vs %s\n", static_cast<uint32_t>(first + i), static_cast<int>(nl - p), p, ref);
                }
                p = nl + 1;
            }
            const int64_t rows = smart_csv_parse_f32(text.data(), n, 1, &wanted, 1, back.data(), block, 1);
            if (rows != static_cast<int64_t>(block)) bad_parse.fetch_add(1);
            p = text.data();
            for (uint64_t i = 0; i < block && rows == static_cast<int64_t>(block); ++i) {
                char *end;
                const float want = static_cast<float>(strtod(p, &end));
                if (memcmp(&want, &back[i], 4) != 0 && !(want != want && back[i] != back[i])) {
                    if (bad_parse.fetch_add(1) < 10) fprintf(stderr, "parse %08x\n", static_cast<uint32_t>(first + i));
                }
                p = static_cast<const char *>(memchr(p, '\n', text.data() + n - p)) + 1;
            }
            checked.fetch_add(block);
        }
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; ++w) pool.emplace_back(work);
    for (auto &t : pool) t.join();
    printf("float32 bit patterns checked: %llu\nformat mismatches against the C library's %%.6e: %llu\nparse mismatches against strtod: %llu\n",
           static_cast<unsigned long long>(checked.load()), static_cast<unsigned long long>(bad_format.load()),
           static_cast<unsigned long long>(bad_parse.load()));
    return bad_format.load() || bad_parse.load() ? 1 : 0;
}
