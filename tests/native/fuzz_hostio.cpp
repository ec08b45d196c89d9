// Mutation fuzzer for the native host decoders (MetaImage, DICOM) and the PNG encoder of libspine_b200's host stage.
// Built by tests/test_hostio.py with -fsanitize=address,undefined together with svb_hostio.cpp itself (no CUDA needed):
// any out-of-bounds access, overflow or leak on a corrupt file aborts the process, which fails the test.
//
//   fuzz_hostio <work_dir> <iterations> <seed file>...
//
// Each seed file (.mha / .mhd / .dcm, written by spine_vision_b200.synthetic) is decoded unchanged first (must succeed), then
// `iterations` mutants (truncation, byte flips, 32-bit fields overwritten with extreme values, ASCII digits rewritten) go
// through header parse + decode with a destination buffer sized from the parsed header -- exactly what hostio.py does.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spine_b200.h"

namespace svb {
thread_local char g_err[512];
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace svb
extern "C" const char* svb_last_error(void) { return svb::g_err; }

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

static std::vector<uint8_t> slurp(const std::string& p) {
    std::vector<uint8_t> b;
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) return b;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    b.resize(n > 0 ? n : 0);
    if (n > 0 && fread(b.data(), 1, n, f) != (size_t)n) b.clear();
    fclose(f);
    return b;
}
static void spit(const std::string& p, const std::vector<uint8_t>& b) {
    FILE* f = fopen(p.c_str(), "wb");
    if (!f) { perror(p.c_str()); exit(2); }
    if (!b.empty()) fwrite(b.data(), 1, b.size(), f);
    fclose(f);
}
static bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

static const size_t MAX_ELEMS = 1u << 22;  // a mutant may claim any size; the caller (hostio.py) allocates what the header says

static int decode_mha(const std::string& path) {
    svb_mha_info info;
    int rc = svb_mha_read_header(path.c_str(), &info);
    if (rc != 0) return rc;
    const double n = (double)info.dim[0] * info.dim[1] * info.dim[2];
    if (!(n > 0) || n > (double)MAX_ELEMS) return -100;
    std::vector<float> dst((size_t)n);
    rc = svb_mha_read_f32(path.c_str(), &info, dst.data(), dst.size());
    {  // the slab entry (the dataset driver's partial decode): random slice ranges, valid and invalid
        const char* paths[1] = {path.c_str()};
        float* dsts[1] = {dst.data()};
        size_t sizes[1] = {dst.size()};
        for (int t = 0; t < 3; ++t) {
            int32_t z0[1] = {(int32_t)(rnd() % (uint32_t)(info.dim[2] + 2)) - 1};
            int32_t z1[1] = {z0[0] + (int32_t)(rnd() % 3)};
            int32_t rcs[1] = {0};
            svb_mha_read_batch_slab_f32(paths, 1, &info, dsts, sizes, z0, z1, 1, rcs);
            if (rcs[0] == 0 && (z0[0] < 0 || z1[0] > info.dim[2] || z0[0] >= z1[0])) { fprintf(stderr, "bad slab accepted\n"); exit(3); }
        }
    }
    if (rc == 0 && dst.size() > 1) {  // a shorter destination must be refused, not overrun
        std::vector<float> small(dst.size() - 1);
        if (svb_mha_read_f32(path.c_str(), &info, small.data(), small.size()) == 0) { fprintf(stderr, "short buffer accepted\n"); exit(3); }
    }
    return rc;
}
static int decode_dcm(const std::string& path) {
    const char* paths[1] = {path.c_str()};
    svb_dicom_info info;
    int32_t rc1 = 0;
    svb_dicom_read_headers(paths, 1, &info, 1, &rc1);
    if (rc1 != 0) return rc1;
    const double n = (double)info.rows * info.cols;
    if (!(n > 0) || n > (double)MAX_ELEMS) return -100;
    std::vector<float> dst((size_t)n);
    float* dsts[1] = {dst.data()};
    size_t sizes[1] = {dst.size()};
    svb_dicom_read_slices_f32(paths, 1, &info, dsts, sizes, 1, &rc1);
    return rc1;
}
static int decode(const std::string& path) { return ends_with(path, ".dcm") ? decode_dcm(path) : decode_mha(path); }

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: fuzz_hostio <work_dir> <iterations> <seed>...\n"); return 2; }
    const std::string work = argv[1];
    const int iters = atoi(argv[2]);
    long ok = 0, bad = 0;
    for (int s = 3; s < argc; ++s) {
        const std::string seed = argv[s];
        const std::vector<uint8_t> blob = slurp(seed);
        if (blob.empty()) { fprintf(stderr, "cannot read %s\n", seed.c_str()); return 2; }
        const char* ext = ends_with(seed, ".dcm") ? ".dcm" : ends_with(seed, ".mhd") ? ".mhd" : ".mha";
        if (decode(seed) != 0) { fprintf(stderr, "seed %s does not decode: %s\n", seed.c_str(), svb::g_err); return 4; }
        std::string mutant = work + "/mutant" + ext;
        if (ends_with(seed, ".mhd")) {  // keep the separate data file reachable under the name the header gives
            svb_mha_info info;
            svb_mha_read_header(seed.c_str(), &info);
            const std::string data = info.data_file;
            const size_t slash = data.find_last_of('/');
            spit(work + "/" + (slash == std::string::npos ? data : data.substr(slash + 1)), slurp(data));
        }
        for (int it = 0; it < iters; ++it) {
            std::vector<uint8_t> m = blob;
            switch (rnd() % 5) {
                case 0: m.resize(rnd() % (m.size() + 1)); break;                                  // truncate
                case 1: for (int k = 0, n = 1 + rnd() % 4; k < n; ++k) m[rnd() % m.size()] = (uint8_t)rnd(); break;  // byte flips
                case 2: {                                                                         // extreme 32-bit field
                    static const uint32_t vals[] = {0u, 1u, 0x7FFFFFFFu, 0x80000000u, 0xFFFFFFFFu, 0xFFFFFFFEu, 0x00010000u};
                    if (m.size() >= 4) { const uint32_t v = vals[rnd() % 7]; memcpy(&m[rnd() % (m.size() - 3)], &v, 4); }
                } break;
                case 3: {                                                                         // rewrite an ASCII digit run in the head
                    const size_t lim = m.size() < 2048 ? m.size() : 2048;
                    for (int tries = 0; tries < 64; ++tries) {
                        const size_t p = rnd() % lim;
                        if (m[p] >= '0' && m[p] <= '9') { static const char repl[] = "0-9.eE+ \\x"; m[p] = (uint8_t)repl[rnd() % 10]; break; }
                    }
                } break;
                default: {                                                                        // drop or double a chunk
                    const size_t a = rnd() % m.size(), n = 1 + rnd() % 64;
                    if (rnd() & 1) m.erase(m.begin() + a, m.begin() + (a + n < m.size() ? a + n : m.size()));
                    else m.insert(m.begin() + a, m.begin() + a, m.begin() + (a + n < m.size() ? a + n : m.size()));
                }
            }
            spit(mutant, m);
            (decode(mutant) == 0 ? ok : bad)++;
        }
    }
    // PNG encoder: odd shapes, capacity exactly at / below the bound
    for (int it = 0; it < 64; ++it) {
        const int h = 1 + rnd() % 97, w = 1 + rnd() % 131;
        std::vector<uint8_t> img((size_t)h * w);
        for (auto& v : img) v = (uint8_t)(rnd() % (it % 3 == 0 ? 2 : 256));
        const size_t cap = svb_png_bound(h, w);
        std::vector<uint8_t> out(cap);
        size_t n = 0;
        if (svb_png_encode_gray8(img.data(), h, w, 6, out.data(), cap, &n) != 0 || n == 0 || n > cap) { fprintf(stderr, "png encode failed\n"); return 5; }
        std::vector<uint8_t> tiny(16);
        if (svb_png_encode_gray8(img.data(), h, w, 6, tiny.data(), tiny.size(), &n) == 0) { fprintf(stderr, "png: tiny buffer accepted\n"); return 5; }
    }
    printf("fuzz_hostio: %ld mutants decoded, %ld rejected, no sanitizer report\n", ok, bad);
    return 0;
}
