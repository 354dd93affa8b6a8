// Mutation fuzzer for the TGA / PNG decoders of the asset pipeline (meteoros_b200/csrc/mt_assets.cpp), meant to run under
// AddressSanitizer + UBSan.  Seeds: any small valid files named a.png b.png c.png d.png e.tga f.tga g.tga in /tmp/seeds
// (tests/test_assets.py::test_decoders_reject_malformed_files builds the same kinds of seeds in memory).
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all tools/fuzz_assets.cpp \
//       meteoros_b200/csrc/mt_assets.cpp -o /tmp/fuzz_assets && /tmp/fuzz_assets 5000000
// Round 1: 5 000 000 mutated inputs, no report (a first plain run had aborted on std::bad_alloc from a mutated size field;
// dimensions and allocations are bounded since, and no exception crosses the C boundary).
#include "../include/meteoros_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
static std::vector<uint8_t> slurp(const char* p) { FILE* f = fopen(p, "rb"); std::vector<uint8_t> v; if (!f) return v; fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); v.resize(n); fread(v.data(), 1, n, f); fclose(f); return v; }
int main(int argc, char** argv)
{
    const char* names[] = { "a.png", "b.png", "c.png", "d.png", "e.tga", "f.tga", "g.tga" };
    std::vector<std::vector<uint8_t>> seeds;
    for (auto n : names) seeds.push_back(slurp((std::string("/tmp/seeds/") + n).c_str()));
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    long iters = argc > 1 ? atol(argv[1]) : 100000, okc = 0;
    std::vector<uint8_t> out(1 << 22);
    for (long it = 0; it < iters; ++it) {
        const size_t k = rnd() % seeds.size();
        std::vector<uint8_t> b = seeds[k];
        const int is_png = k < 4;
        switch (rnd() % 5) {
        case 0: for (int j = 0, m = 1 + rnd() % 6; j < m; ++j) b[rnd() % b.size()] = (uint8_t)rnd(); break;
        case 1: b.resize(rnd() % (b.size() + 1)); break;
        case 2: { size_t i = rnd() % b.size(); b.insert(b.begin() + i, (uint8_t)rnd()); } break;
        case 3: { size_t i = rnd() % (b.size() > 4 ? b.size() - 4 : 1); for (int j = 0; j < 4 && i + j < b.size(); ++j) b[i + j] = (uint8_t)rnd(); } break;
        default: { size_t i = rnd() % b.size(); b[i] ^= (uint8_t)(1u << (rnd() % 8)); }
        }
        uint32_t w = 0, h = 0;
        if (mtxDecodeImage(b.data(), b.size(), is_png, nullptr, 0, &w, &h) == MT_OK && (size_t)w * h * 4 <= out.size())
            okc += mtxDecodeImage(b.data(), b.size(), is_png, out.data(), out.size(), &w, &h) == MT_OK;
    }
    printf("iterations %ld, decoded %ld\n", iters, okc);
    return 0;
}
