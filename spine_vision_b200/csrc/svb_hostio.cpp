// svb_hostio.cpp -- host input / output stage either side of the GPU path.
// Plain C++ (no CUDA): a MetaImage (.mha / .mhd) decoder that fills caller-owned (pinned) float32 buffers, and an 8-bit
// greyscale PNG encoder, both fanned out over a small thread pool so that the host side keeps up with the GPU.
//
//   reference                                                        here
//   read_medical_image -> read_mha -> sitk.ReadImage                 svb_mha_read_header / svb_mha_read_f32(_batch)
//     (io/readers.py:65-73, 128-161; spider.py:115)
//   Image.fromarray(crop).save(path)   (PNG "L", zlib level 6)       svb_png_encode_gray8 / svb_png_write_gray8_batch
//     (datasets/classification/spider.py:158, phenikaa.py:213)
//
// PNG parity is on the DECODED pixels (the file bytes depend on the encoder's filter heuristics and are not part of the
// reference's contract: ClassificationDataset re-opens the files with PIL, training/datasets/classification.py:288-294).
#include <zlib.h>

#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/spine_b200.h"

namespace svb {
int set_error(int code, const char* fmt, ...);  // svb_common.cu (thread-local message)
}
using svb::set_error;

namespace {

// ------------------------------------------------------------------------------------------------ thread fan-out
template <typename F>
void parallel_for(int n, int n_threads, F&& fn) {
    if (n <= 0) return;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > n) nt = n;
    if (nt == 1) {
        for (int i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&] {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto& th : pool) th.join();
}

// No C++ exception may leave an extern "C" entry or a worker thread (std::terminate would take the whole process down):
// allocation failures and anything else thrown by the body become an error code + message.
template <typename F>
int guarded(const char* what, F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return set_error(SVB_ERR_IO, "%s: out of host memory", what);
    } catch (const std::exception& e) {
        return set_error(SVB_ERR_IO, "%s: %s", what, e.what());
    } catch (...) {
        return set_error(SVB_ERR_IO, "%s: unknown C++ exception", what);
    }
}

// ------------------------------------------------------------------------------------------------ PNG
void put_be32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}
// one chunk: length, type, data, crc(type + data); returns bytes written
size_t put_chunk(uint8_t* out, const char type[4], const uint8_t* data, uint32_t len) {
    put_be32(out, len);
    memcpy(out + 4, type, 4);
    if (len) memcpy(out + 8, data, len);
    uint32_t crc = (uint32_t)crc32(0L, out + 4, 4 + len);
    put_be32(out + 8 + len, crc);
    return 12 + (size_t)len;
}
// PNG filter types 0-4 for one greyscale row (bpp = 1); picks the one with the least sum of |signed residual|
// (the heuristic the PNG specification recommends); writes [type byte][w residuals] to dst
void filter_row(const uint8_t* cur, const uint8_t* prev /* may be null */, int w, uint8_t* dst, std::vector<uint8_t>& scratch) {
    // left / up / up-left neighbours as padded rows, so that every filter is one branch-free loop the compiler vectorises
    scratch.resize((size_t)3 * (w + 1) + (size_t)4 * w);
    uint8_t* a = scratch.data();          // a[x] = cur[x-1]
    uint8_t* b = a + (w + 1);             // b[x] = prev[x]
    uint8_t* c = b + (w + 1);             // c[x] = prev[x-1]
    uint8_t* f[5] = {nullptr, c + (w + 1), c + (w + 1) + w, c + (w + 1) + 2 * (size_t)w, c + (w + 1) + 3 * (size_t)w};
    a[0] = 0;
    memcpy(a + 1, cur, (size_t)w - 1);
    if (prev) {
        memcpy(b, prev, (size_t)w);
        c[0] = 0;
        memcpy(c + 1, prev, (size_t)w - 1);
    } else {
        memset(b, 0, (size_t)w);
        memset(c, 0, (size_t)w);
    }
    uint32_t sum[5] = {0, 0, 0, 0, 0};
    for (int x = 0; x < w; ++x) sum[0] += (uint32_t)abs((int)(int8_t)cur[x]);
    for (int x = 0; x < w; ++x) { const uint8_t r = (uint8_t)(cur[x] - a[x]); f[1][x] = r; sum[1] += (uint32_t)abs((int)(int8_t)r); }
    for (int x = 0; x < w; ++x) { const uint8_t r = (uint8_t)(cur[x] - b[x]); f[2][x] = r; sum[2] += (uint32_t)abs((int)(int8_t)r); }
    for (int x = 0; x < w; ++x) { const uint8_t r = (uint8_t)(cur[x] - ((a[x] + b[x]) >> 1)); f[3][x] = r; sum[3] += (uint32_t)abs((int)(int8_t)r); }
    for (int x = 0; x < w; ++x) {
        const int aa = a[x], bb = b[x], cc = c[x];
        const int p = aa + bb - cc, pa = abs(p - aa), pb = abs(p - bb), pc = abs(p - cc);
        const int pred = (pa <= pb && pa <= pc) ? aa : (pb <= pc ? bb : cc);
        const uint8_t r = (uint8_t)(cur[x] - pred);
        f[4][x] = r;
        sum[4] += (uint32_t)abs((int)(int8_t)r);
    }
    int best = 0;
    for (int t = 1; t < 5; ++t)
        if (sum[t] < sum[best]) best = t;
    dst[0] = (uint8_t)best;
    memcpy(dst + 1, best == 0 ? cur : f[best], (size_t)w);
}

int png_encode(const uint8_t* img, int h, int w, int level, uint8_t* out, size_t cap, size_t* out_len) {
    if (!img || !out || !out_len || h <= 0 || w <= 0) return set_error(SVB_ERR_INVALID_ARG, "png: bad arguments (h=%d w=%d)", h, w);
    if (level < 0 || level > 9) level = 6;
    const size_t raw_len = (size_t)h * ((size_t)w + 1);
    std::vector<uint8_t> raw(raw_len), scratch;
    for (int y = 0; y < h; ++y)
        filter_row(img + (size_t)y * w, y ? img + (size_t)(y - 1) * w : nullptr, w, raw.data() + (size_t)y * (w + 1), scratch);
    uLongf zlen = compressBound((uLong)raw_len);
    std::vector<uint8_t> z(zlen);
    const int zr = compress2(z.data(), &zlen, raw.data(), (uLong)raw_len, level);
    if (zr != Z_OK) return set_error(SVB_ERR_IO, "png: zlib compress2 failed (%d)", zr);
    const size_t need = 8 + (12 + 13) + (12 + (size_t)zlen) + 12;
    if (cap < need) return set_error(SVB_ERR_WORKSPACE_TOO_SMALL, "png: output buffer %zu < %zu bytes", cap, need);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    size_t o = 0;
    memcpy(out, sig, 8);
    o += 8;
    uint8_t ihdr[13];
    put_be32(ihdr, (uint32_t)w);
    put_be32(ihdr + 4, (uint32_t)h);
    ihdr[8] = 8;   // bit depth
    ihdr[9] = 0;   // colour type 0 = greyscale (PIL mode "L")
    ihdr[10] = 0;  // deflate
    ihdr[11] = 0;  // adaptive filtering
    ihdr[12] = 0;  // no interlace
    o += put_chunk(out + o, "IHDR", ihdr, 13);
    o += put_chunk(out + o, "IDAT", z.data(), (uint32_t)zlen);
    o += put_chunk(out + o, "IEND", nullptr, 0);
    *out_len = o;
    return SVB_OK;
}

int write_file(const char* path, const uint8_t* data, size_t len) {
    FILE* f = fopen(path, "wb");
    if (!f) return set_error(SVB_ERR_IO, "cannot open %s for writing: %s", path, strerror(errno));
    const size_t n = fwrite(data, 1, len, f);
    const int rc = fclose(f);
    if (n != len || rc != 0) return set_error(SVB_ERR_IO, "short write to %s", path);
    return SVB_OK;
}

// ------------------------------------------------------------------------------------------------ MetaImage
std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) ++a;
    while (b > a && isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
bool truthy(const std::string& v) { return !v.empty() && (v[0] == 'T' || v[0] == 't' || v[0] == '1'); }
int parse_doubles(const std::string& v, double* out, int max_n) {
    int n = 0;
    const char* p = v.c_str();
    while (n < max_n) {
        char* e = nullptr;
        const double d = strtod(p, &e);
        if (e == p) break;
        out[n++] = d;
        p = e;
    }
    return n;
}
struct MetType { const char* name; int id; int bytes; };
const MetType kTypes[] = {
    {"MET_CHAR", SVB_MHA_I8, 1},     {"MET_UCHAR", SVB_MHA_U8, 1},   {"MET_SHORT", SVB_MHA_I16, 2},  {"MET_USHORT", SVB_MHA_U16, 2},
    {"MET_INT", SVB_MHA_I32, 4},     {"MET_UINT", SVB_MHA_U32, 4},   {"MET_LONG", SVB_MHA_I32, 4},   {"MET_ULONG", SVB_MHA_U32, 4},
    {"MET_LONG_LONG", SVB_MHA_I64, 8}, {"MET_ULONG_LONG", SVB_MHA_U64, 8}, {"MET_FLOAT", SVB_MHA_F32, 4}, {"MET_DOUBLE", SVB_MHA_F64, 8},
};

int mha_header(const char* path, svb_mha_info* info) {
    if (!path || !info) return set_error(SVB_ERR_INVALID_ARG, "mha: null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(SVB_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    memset(info, 0, sizeof(*info));
    info->ndim = 0;
    info->channels = 1;
    info->data_offset = -1;
    info->header_size = 0;
    for (int i = 0; i < 3; ++i) { info->dim[i] = 1; info->spacing[i] = 1.0; info->direction[4 * i] = 1.0; }
    bool have_type = false, have_dims = false, have_file = false, binary = true;
    double tm[9];
    int n_tm = 0;
    std::string line;
    for (;;) {
        line.clear();
        int c;
        while ((c = fgetc(f)) != EOF && c != '\n') {
            line.push_back((char)c);
            if (line.size() > 8192) { fclose(f); return set_error(SVB_ERR_FORMAT, "%s: header line too long (not a MetaImage?)", path); }
        }
        if (line.empty() && c == EOF) break;
        const size_t eq = line.find('=');
        if (eq == std::string::npos) {
            if (trim(line).empty()) { if (c == EOF) break; continue; }
            fclose(f);
            return set_error(SVB_ERR_FORMAT, "%s: malformed header line '%s'", path, line.substr(0, 60).c_str());
        }
        const std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
        if (key == "NDims") info->ndim = atoi(val.c_str());
        else if (key == "DimSize") {
            double d[3] = {1, 1, 1};
            const int n = parse_doubles(val, d, 3);
            for (int i = 0; i < 3; ++i) {
                // checked BEFORE the cast: a double outside int32 is undefined behaviour to convert, and a corrupt header must
                // be a format error for this one series (spider.py:131-133 skips it), never a multi-terabyte allocation
                if (i < n && !(d[i] >= 1.0 && d[i] <= 1048576.0)) { fclose(f); return set_error(SVB_ERR_FORMAT, "%s: DimSize[%d] = %g is outside 1 .. 2^20", path, i, d[i]); }
                info->dim[i] = i < n ? (int32_t)d[i] : 1;
            }
            have_dims = n > 0;
        } else if (key == "ElementSpacing" || key == "ElementSize") {
            // ElementSpacing wins when both are present (MetaIO)
            if (key == "ElementSpacing" || !info->has_spacing) {
                double d[3] = {1, 1, 1};
                parse_doubles(val, d, 3);
                for (int i = 0; i < 3; ++i) info->spacing[i] = d[i];
                if (key == "ElementSpacing") info->has_spacing = 1;
            }
        } else if (key == "Offset" || key == "Position" || key == "Origin") parse_doubles(val, info->origin, 3);
        else if (key == "TransformMatrix" || key == "Orientation" || key == "Rotation") n_tm = parse_doubles(val, tm, 9);
        else if (key == "ElementNumberOfChannels") info->channels = atoi(val.c_str());
        else if (key == "BinaryData") binary = truthy(val);
        else if (key == "BinaryDataByteOrderMSB" || key == "ElementByteOrderMSB") info->big_endian = truthy(val);
        else if (key == "CompressedData") info->compressed = truthy(val);
        else if (key == "CompressedDataSize") info->compressed_size = atoll(val.c_str());
        else if (key == "HeaderSize") info->header_size = atoll(val.c_str());
        else if (key == "ElementType") {
            for (const MetType& t : kTypes)
                if (val == t.name) { info->element_type = t.id; info->element_bytes = t.bytes; have_type = true; }
            if (!have_type) { fclose(f); return set_error(SVB_ERR_FORMAT, "%s: unsupported ElementType %s", path, val.c_str()); }
        } else if (key == "ElementDataFile") {
            have_file = true;
            if (val == "LOCAL") info->data_offset = (int64_t)ftell(f);
            else {
                if (val.size() >= sizeof(info->data_file)) { fclose(f); return set_error(SVB_ERR_FORMAT, "%s: data file name too long", path); }
                if (val == "LIST" || val.find('%') != std::string::npos) { fclose(f); return set_error(SVB_ERR_FORMAT, "%s: ElementDataFile = %s (per-slice files) is not supported", path, val.c_str()); }
                std::string p(path);
                const size_t slash = p.find_last_of('/');
                const std::string full = (val[0] == '/' || slash == std::string::npos) ? val : p.substr(0, slash + 1) + val;
                if (full.size() >= sizeof(info->data_file)) { fclose(f); return set_error(SVB_ERR_FORMAT, "%s: data file path too long", path); }
                strcpy(info->data_file, full.c_str());
            }
            break;  // ElementDataFile is the last header field by definition
        }
    }
    fclose(f);
    if (!have_type || !have_dims || !have_file) return set_error(SVB_ERR_FORMAT, "%s: not a MetaImage header (DimSize / ElementType / ElementDataFile missing)", path);
    if (!binary) return set_error(SVB_ERR_FORMAT, "%s: ASCII MetaImage data is not supported", path);
    if (info->ndim < 2 || info->ndim > 3) return set_error(SVB_ERR_FORMAT, "%s: NDims = %d (2 or 3 supported)", path, info->ndim);
    if (info->channels != 1) return set_error(SVB_ERR_FORMAT, "%s: %d channels (scalar images only)", path, info->channels);
    for (int i = 0; i < 3; ++i)
        if (info->dim[i] <= 0) return set_error(SVB_ERR_FORMAT, "%s: bad DimSize", path);
    {
        // the voxel count must be backed by the data that is actually there: raw data by the file's bytes, a zlib stream by at
        // most 1032x its compressed size (deflate's maximum expansion)
        const double voxels = (double)info->dim[0] * info->dim[1] * info->dim[2];
        if (voxels > 8589934592.0) return set_error(SVB_ERR_FORMAT, "%s: DimSize product %.0f exceeds 2^33 voxels", path, voxels);
        const double all_bytes = voxels * info->element_bytes;
        const char* data_path = info->data_offset >= 0 ? path : info->data_file;
        FILE* df = fopen(data_path, "rb");
        if (df) {
            fseek(df, 0, SEEK_END);
            const double file_bytes = (double)ftell(df);
            fclose(df);
            const double start = info->data_offset >= 0 ? (double)info->data_offset : (info->header_size > 0 ? (double)info->header_size : 0.0);
            const double avail = info->compressed && info->compressed_size > 0 ? (double)info->compressed_size : file_bytes - start;
            const double limit = info->compressed ? avail * 1032.0 + 65536.0 : avail;
            if (avail < 0 || all_bytes > limit)
                return set_error(SVB_ERR_FORMAT, "%s: DimSize asks for %.0f bytes of voxels, the data holds at most %.0f", path, all_bytes, limit);
        }  // a missing data file is reported by the read, with its own message
    }
    // ITK: row i of TransformMatrix is the direction cosine of image axis i, i.e. COLUMN i of image.GetDirection()
    const int nd = info->ndim;
    if (n_tm >= nd * nd) {
        for (int i = 0; i < 9; ++i) info->direction[i] = (i % 4 == 0) ? 1.0 : 0.0;
        for (int i = 0; i < nd; ++i)
            for (int j = 0; j < nd; ++j) info->direction[j * 3 + i] = tm[i * nd + j];
    }
    return SVB_OK;
}

template <typename S>
void convert(const uint8_t* src, size_t n, bool swap, float* dst) {
    for (size_t i = 0; i < n; ++i) {
        uint8_t b[sizeof(S)];
        memcpy(b, src + i * sizeof(S), sizeof(S));
        if (swap)
            for (size_t k = 0; k < sizeof(S) / 2; ++k) { const uint8_t t = b[k]; b[k] = b[sizeof(S) - 1 - k]; b[sizeof(S) - 1 - k] = t; }
        S v;
        memcpy(&v, b, sizeof(S));
        dst[i] = static_cast<float>(v);
    }
}

// Inflate `want` bytes of a zlib stream starting `skip` bytes into the inflated data (a slab of z slices): the bytes in front are
// inflated into a scratch block and dropped, the stream is not read past the slab.
int inflate_range(const uint8_t* z, size_t zbytes, size_t skip, uint8_t* out, size_t want) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) return Z_MEM_ERROR;
    zs.next_in = const_cast<Bytef*>(z);
    zs.avail_in = (uInt)zbytes;
    std::vector<uint8_t> scratch(skip ? (size_t)1 << 16 : 0);
    int zr = Z_OK;
    while (skip > 0 && zr == Z_OK) {
        const size_t chunk = skip < scratch.size() ? skip : scratch.size();
        zs.next_out = scratch.data();
        zs.avail_out = (uInt)chunk;
        zr = inflate(&zs, Z_NO_FLUSH);
        skip -= chunk - zs.avail_out;
        if (zr == Z_OK && zs.avail_out == chunk && zs.avail_in == 0) zr = Z_BUF_ERROR;  // no progress: truncated stream
    }
    if (skip == 0 && (zr == Z_OK || (zr == Z_STREAM_END && want == 0))) {
        zs.next_out = out;
        zs.avail_out = (uInt)want;
        while (zs.avail_out > 0 && zr == Z_OK) {
            const uInt before = zs.avail_out;
            zr = inflate(&zs, Z_NO_FLUSH);
            if (zr == Z_OK && zs.avail_out == before && zs.avail_in == 0) zr = Z_BUF_ERROR;
        }
        if (zs.avail_out == 0 && (zr == Z_OK || zr == Z_STREAM_END)) zr = Z_OK;
        else if (zr == Z_OK || zr == Z_STREAM_END) zr = Z_BUF_ERROR;  // the stream ended inside the slab
    } else if (zr == Z_OK || zr == Z_STREAM_END) {
        zr = Z_BUF_ERROR;  // the stream ended in front of the slab
    }
    inflateEnd(&zs);
    return zr;
}

// Slices [z0, z1) of the volume into dst + z0 * (slice elements); dst is the buffer of the WHOLE volume (what lies outside the
// slab is not touched).  z0 = 0, z1 = dim[2] is the whole volume.
int mha_read_slab(const char* path, const svb_mha_info* info, float* dst, size_t dst_elems, int z0, int z1) {
    if (!path || !info || !dst) return set_error(SVB_ERR_INVALID_ARG, "mha: null argument");
    const size_t n_all = (size_t)info->dim[0] * info->dim[1] * info->dim[2];
    if (dst_elems < n_all) return set_error(SVB_ERR_WORKSPACE_TOO_SMALL, "mha: destination holds %zu elements, volume has %zu", dst_elems, n_all);
    if (z0 < 0 || z1 > info->dim[2] || z0 >= z1) return set_error(SVB_ERR_INVALID_ARG, "mha: slab [%d, %d) outside the %d slices", z0, z1, info->dim[2]);
    const bool whole = z0 == 0 && z1 == info->dim[2];
    const size_t slice_elems = (size_t)info->dim[0] * info->dim[1];
    const size_t n = slice_elems * (size_t)(z1 - z0);
    const size_t skip_bytes = slice_elems * (size_t)z0 * (size_t)info->element_bytes;
    const size_t all_bytes = n_all * (size_t)info->element_bytes;
    dst += slice_elems * (size_t)z0;
    const size_t raw_bytes = n * (size_t)info->element_bytes;
    const char* data_path = info->data_offset >= 0 ? path : info->data_file;
    FILE* f = fopen(data_path, "rb");
    if (!f) return set_error(SVB_ERR_IO, "cannot open %s: %s", data_path, strerror(errno));
    std::vector<uint8_t> raw(raw_bytes);
    int rc = SVB_OK;
    if (info->compressed && !whole) {
        int64_t start = info->data_offset >= 0 ? info->data_offset : (info->header_size > 0 ? info->header_size : 0);
        fseek(f, 0, SEEK_END);
        const int64_t end = (int64_t)ftell(f);
        int64_t zbytes = info->compressed_size > 0 ? info->compressed_size : end - start;
        if (zbytes <= 0 || start + zbytes > end) { fclose(f); return set_error(SVB_ERR_FORMAT, "%s: compressed data truncated", data_path); }
        std::vector<uint8_t> z((size_t)zbytes);
        fseek(f, (long)start, SEEK_SET);
        if (fread(z.data(), 1, (size_t)zbytes, f) != (size_t)zbytes) { fclose(f); return set_error(SVB_ERR_IO, "%s: short read", data_path); }
        const int zr = inflate_range(z.data(), (size_t)zbytes, skip_bytes, raw.data(), raw_bytes);
        if (zr != Z_OK) rc = set_error(SVB_ERR_FORMAT, "%s: zlib inflate of slices [%d, %d) failed (%d)", data_path, z0, z1, zr);
    } else if (info->compressed) {
        // the compressed stream runs to CompressedDataSize, or to the end of the file when the field is absent
        int64_t start = info->data_offset >= 0 ? info->data_offset : (info->header_size > 0 ? info->header_size : 0);
        fseek(f, 0, SEEK_END);
        const int64_t end = (int64_t)ftell(f);
        int64_t zbytes = info->compressed_size > 0 ? info->compressed_size : end - start;
        if (zbytes <= 0 || start + zbytes > end) { fclose(f); return set_error(SVB_ERR_FORMAT, "%s: compressed data truncated", data_path); }
        std::vector<uint8_t> z((size_t)zbytes);
        fseek(f, (long)start, SEEK_SET);
        if (fread(z.data(), 1, (size_t)zbytes, f) != (size_t)zbytes) { fclose(f); return set_error(SVB_ERR_IO, "%s: short read", data_path); }
        uLongf out_len = (uLongf)raw_bytes;
        const int zr = uncompress(raw.data(), &out_len, z.data(), (uLong)zbytes);
        if (zr != Z_OK || out_len != raw_bytes) rc = set_error(SVB_ERR_FORMAT, "%s: zlib inflate failed (%d) or size mismatch (%lu of %zu bytes)", data_path, zr, (unsigned long)out_len, raw_bytes);
    } else {
        int64_t start;
        if (info->data_offset >= 0) start = info->data_offset;
        else if (info->header_size == -1) {  // MetaIO: data sits at the end of the file
            fseek(f, 0, SEEK_END);
            start = (int64_t)ftell(f) - (int64_t)all_bytes;
        } else start = info->header_size;
        if (start >= 0) start += (int64_t)skip_bytes;
        if (start < 0 || fseek(f, (long)start, SEEK_SET) != 0 || fread(raw.data(), 1, raw_bytes, f) != raw_bytes)
            rc = set_error(SVB_ERR_FORMAT, "%s: raw data truncated (%zu bytes expected)", data_path, raw_bytes);
    }
    fclose(f);
    if (rc) return rc;
    const uint16_t probe = 1;
    const bool host_big = *reinterpret_cast<const uint8_t*>(&probe) == 0;
    const bool swap = (info->big_endian != 0) != host_big && info->element_bytes > 1;
    switch (info->element_type) {
        case SVB_MHA_I8: convert<int8_t>(raw.data(), n, false, dst); break;
        case SVB_MHA_U8: convert<uint8_t>(raw.data(), n, false, dst); break;
        case SVB_MHA_I16: convert<int16_t>(raw.data(), n, swap, dst); break;
        case SVB_MHA_U16: convert<uint16_t>(raw.data(), n, swap, dst); break;
        case SVB_MHA_I32: convert<int32_t>(raw.data(), n, swap, dst); break;
        case SVB_MHA_U32: convert<uint32_t>(raw.data(), n, swap, dst); break;
        case SVB_MHA_I64: convert<int64_t>(raw.data(), n, swap, dst); break;
        case SVB_MHA_U64: convert<uint64_t>(raw.data(), n, swap, dst); break;
        case SVB_MHA_F32: convert<float>(raw.data(), n, swap, dst); break;
        case SVB_MHA_F64: convert<double>(raw.data(), n, swap, dst); break;
        default: return set_error(SVB_ERR_FORMAT, "%s: unknown element type %d", path, info->element_type);
    }
    return SVB_OK;
}

int mha_read(const char* path, const svb_mha_info* info, float* dst, size_t dst_elems) {
    if (!info) return set_error(SVB_ERR_INVALID_ARG, "mha: null argument");
    return mha_read_slab(path, info, dst, dst_elems, 0, info->dim[2]);
}


// ------------------------------------------------------------------------------------------------ DICOM (one slice per file)
// The subset the Phenikaa series need (io/readers.py:48-73 -> sitk.ImageSeriesReader over GDCM): Part-10 files, implicit
// or explicit VR little endian (and explicit big endian), native (uncompressed) pixel data, 8/16/32-bit monochrome.
struct DcmCursor {
    const uint8_t* p;
    size_t n, pos;
    bool explicit_vr, big;
    uint16_t u16() {
        const uint16_t v = big ? (uint16_t)((p[pos] << 8) | p[pos + 1]) : (uint16_t)(p[pos] | (p[pos + 1] << 8));
        pos += 2;
        return v;
    }
    uint32_t u32() {
        const uint32_t v = big ? ((uint32_t)p[pos] << 24 | (uint32_t)p[pos + 1] << 16 | (uint32_t)p[pos + 2] << 8 | p[pos + 3])
                               : ((uint32_t)p[pos] | (uint32_t)p[pos + 1] << 8 | (uint32_t)p[pos + 2] << 16 | (uint32_t)p[pos + 3] << 24);
        pos += 4;
        return v;
    }
};
bool long_vr(const char vr[2]) {
    static const char* k[] = {"OB", "OW", "OF", "OD", "OL", "SQ", "UC", "UR", "UT", "UN"};
    for (const char* s : k)
        if (vr[0] == s[0] && vr[1] == s[1]) return true;
    return false;
}
// skip a sequence / item of undefined length: walk nested elements until the matching delimiter
bool dcm_skip_undefined(DcmCursor& c, int depth);
bool dcm_next(DcmCursor& c, uint16_t& g, uint16_t& e, char vr[2], uint32_t& len) {
    if (c.pos + 8 > c.n) return false;
    g = c.u16();
    e = c.u16();
    vr[0] = vr[1] = 0;
    if (g == 0xFFFE) {  // item / delimiters: always implicit form
        len = c.u32();
        return true;
    }
    if (c.explicit_vr) {
        vr[0] = (char)c.p[c.pos];
        vr[1] = (char)c.p[c.pos + 1];
        c.pos += 2;
        if (long_vr(vr)) {
            c.pos += 2;
            if (c.pos + 4 > c.n) return false;
            len = c.u32();
        } else {
            len = c.u16();
        }
    } else {
        len = c.u32();
    }
    return true;
}
bool dcm_skip_undefined(DcmCursor& c, int depth) {
    if (depth > 16) return false;
    for (;;) {
        uint16_t g, e;
        char vr[2];
        uint32_t len;
        if (!dcm_next(c, g, e, vr, len)) return false;
        if (g == 0xFFFE && (e == 0xE0DD || e == 0xE00D)) return true;  // sequence / item delimiter
        if (len == 0xFFFFFFFFu) {
            if (!dcm_skip_undefined(c, depth + 1)) return false;
        } else {
            if (c.pos + len > c.n) return false;
            c.pos += len;
        }
    }
}
std::string dcm_str(const DcmCursor& c, uint32_t len) {
    std::string s(reinterpret_cast<const char*>(c.p + c.pos), len);
    while (!s.empty() && (s.back() == ' ' || s.back() == '\0')) s.pop_back();
    size_t a = 0;
    while (a < s.size() && s[a] == ' ') ++a;
    return s.substr(a);
}
int dcm_numbers(const std::string& s, double* out, int max_n) {  // "a\b\c" decimal strings
    int n = 0;
    size_t pos = 0;
    while (n < max_n && pos <= s.size()) {
        size_t q = s.find('\\', pos);
        if (q == std::string::npos) q = s.size();
        const std::string t = s.substr(pos, q - pos);
        char* e = nullptr;
        const double d = strtod(t.c_str(), &e);
        if (e == t.c_str()) break;
        out[n++] = d;
        pos = q + 1;
    }
    return n;
}
int read_whole(const char* path, std::vector<uint8_t>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(SVB_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    const size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
    fclose(f);
    if (got != buf.size()) return set_error(SVB_ERR_IO, "%s: short read", path);
    return SVB_OK;
}

int dicom_header(const char* path, svb_dicom_info* info, std::vector<uint8_t>* keep) {
    if (!path || !info) return set_error(SVB_ERR_INVALID_ARG, "dicom: null argument");
    std::vector<uint8_t> local;
    std::vector<uint8_t>& buf = keep ? *keep : local;
    if (int rc = read_whole(path, buf)) return rc;
    memset(info, 0, sizeof(*info));
    info->rescale_slope = 1.0;
    info->samples_per_pixel = 1;
    info->pixel_spacing[0] = info->pixel_spacing[1] = 1.0;
    info->orientation[0] = 1.0;
    info->orientation[4] = 1.0;
    info->pixel_offset = -1;
    if (buf.size() < 132 || memcmp(buf.data() + 128, "DICM", 4) != 0)
        return set_error(SVB_ERR_FORMAT, "%s: not a DICOM Part-10 file (no DICM marker)", path);
    DcmCursor c{buf.data(), buf.size(), 132, true, false};
    std::string ts = "1.2.840.10008.1.2.1";
    bool in_meta = true;
    for (;;) {
        const size_t at = c.pos;
        if (in_meta) {  // leaving group 0002 switches to the data set's transfer syntax
            if (c.pos + 2 > c.n) break;
            const uint16_t g0 = (uint16_t)(c.p[c.pos] | (c.p[c.pos + 1] << 8));
            if (g0 != 0x0002) {
                in_meta = false;
                if (ts == "1.2.840.10008.1.2") c.explicit_vr = false;
                else if (ts == "1.2.840.10008.1.2.1") c.explicit_vr = true;
                else if (ts == "1.2.840.10008.1.2.2") { c.explicit_vr = true; c.big = true; }
                else if (ts == "1.2.840.10008.1.2.5") { c.explicit_vr = true; info->encapsulation = SVB_DICOM_RLE; }
                else if (ts == "1.2.840.10008.1.2.4.57" || ts == "1.2.840.10008.1.2.4.70") { c.explicit_vr = true; info->encapsulation = SVB_DICOM_JPEG_LOSSLESS; }
                else return set_error(SVB_ERR_FORMAT, "%s: transfer syntax %s is not supported (native, RLE Lossless and JPEG Lossless are)", path, ts.c_str());
                info->big_endian = c.big ? 1 : 0;
            }
        }
        uint16_t g, e;
        char vr[2];
        uint32_t len;
        if (!dcm_next(c, g, e, vr, len)) break;
        if (g == 0x7FE0 && e == 0x0010) {
            if ((len == 0xFFFFFFFFu) != (info->encapsulation != SVB_DICOM_NATIVE))
                return set_error(SVB_ERR_FORMAT, "%s: pixel data %s but the transfer syntax says %s", path,
                                 len == 0xFFFFFFFFu ? "is encapsulated" : "has a defined length", info->encapsulation ? "encapsulated" : "native");
            info->pixel_offset = (int64_t)c.pos;
            info->pixel_bytes = len == 0xFFFFFFFFu ? (int64_t)(c.n - c.pos) : (int64_t)len;
            if (len != 0xFFFFFFFFu && c.pos + len > c.n) return set_error(SVB_ERR_FORMAT, "%s: pixel data runs past the end of the file", path);
            break;
        }
        if (len == 0xFFFFFFFFu) {
            if (!dcm_skip_undefined(c, 0)) return set_error(SVB_ERR_FORMAT, "%s: malformed sequence at byte %zu", path, at);
            continue;
        }
        if (c.pos + len > c.n) return set_error(SVB_ERR_FORMAT, "%s: element (%04x,%04x) runs past the end of the file", path, g, e);
        const uint32_t tag = ((uint32_t)g << 16) | e;
        double d[6];
        switch (tag) {
            case 0x00020010: ts = dcm_str(c, len); break;
            case 0x0020000E: { const std::string s = dcm_str(c, len); strncpy(info->series_uid, s.c_str(), sizeof(info->series_uid) - 1); } break;
            case 0x00200013: info->instance_number = atoi(dcm_str(c, len).c_str()); break;
            case 0x00200032: if (dcm_numbers(dcm_str(c, len), d, 3) == 3) { memcpy(info->position, d, 24); info->has_position = 1; } break;
            case 0x00200037: if (dcm_numbers(dcm_str(c, len), d, 6) == 6) { memcpy(info->orientation, d, 48); info->has_orientation = 1; } break;
            case 0x00280002: if (len >= 2) { DcmCursor t = c; info->samples_per_pixel = t.u16(); } break;
            case 0x00280004: { const std::string s = dcm_str(c, len); info->monochrome1 = s == "MONOCHROME1"; } break;
            case 0x00280010: if (len >= 2) { DcmCursor t = c; info->rows = t.u16(); } break;
            case 0x00280011: if (len >= 2) { DcmCursor t = c; info->cols = t.u16(); } break;
            case 0x00280030: if (dcm_numbers(dcm_str(c, len), d, 2) == 2) { info->pixel_spacing[0] = d[0]; info->pixel_spacing[1] = d[1]; info->has_spacing = 1; } break;
            case 0x00280100: if (len >= 2) { DcmCursor t = c; info->bits_allocated = t.u16(); } break;
            case 0x00280101: if (len >= 2) { DcmCursor t = c; info->bits_stored = t.u16(); } break;
            case 0x00280103: if (len >= 2) { DcmCursor t = c; info->pixel_representation = t.u16(); } break;
            case 0x00281052: if (dcm_numbers(dcm_str(c, len), d, 1) == 1) info->rescale_intercept = d[0]; break;
            case 0x00281053: if (dcm_numbers(dcm_str(c, len), d, 1) == 1) info->rescale_slope = d[0]; break;
            case 0x00180050: if (dcm_numbers(dcm_str(c, len), d, 1) == 1) info->slice_thickness = d[0]; break;
            case 0x00180088: if (dcm_numbers(dcm_str(c, len), d, 1) == 1) info->spacing_between_slices = d[0]; break;
            default: break;
        }
        c.pos += len;
    }
    if (info->pixel_offset < 0) return set_error(SVB_ERR_FORMAT, "%s: no pixel data element", path);
    if (info->rows <= 0 || info->cols <= 0) return set_error(SVB_ERR_FORMAT, "%s: Rows / Columns missing", path);
    if (info->samples_per_pixel != 1) return set_error(SVB_ERR_FORMAT, "%s: %d samples per pixel (monochrome only)", path, info->samples_per_pixel);
    if (info->bits_allocated != 8 && info->bits_allocated != 16 && info->bits_allocated != 32)
        return set_error(SVB_ERR_FORMAT, "%s: BitsAllocated = %d", path, info->bits_allocated);
    if (info->bits_stored <= 0 || info->bits_stored > info->bits_allocated) info->bits_stored = info->bits_allocated;
    if (info->monochrome1)
        return set_error(SVB_ERR_FORMAT, "%s: PhotometricInterpretation MONOCHROME1 (ITK / GDCM invert it to MONOCHROME2; not decoded here)", path);
    if (info->encapsulation != SVB_DICOM_NATIVE && info->bits_allocated > 16)
        return set_error(SVB_ERR_FORMAT, "%s: encapsulated pixel data with BitsAllocated = %d", path, info->bits_allocated);
    const int64_t need = info->encapsulation != SVB_DICOM_NATIVE ? 8 : (int64_t)info->rows * info->cols * (info->bits_allocated / 8);
    if (info->pixel_bytes < need) return set_error(SVB_ERR_FORMAT, "%s: pixel data holds %lld bytes, %lld expected", path, (long long)info->pixel_bytes, (long long)need);
    return SVB_OK;
}

// stored value -> float: BitsStored < BitsAllocated is masked (unsigned) or sign-extended from the high stored bit (signed),
// as GDCM's "rescale stored bits" step does; then slope / intercept in double like ITK's rescale functor
template <typename S>
void convert_rescaled(const uint8_t* src, size_t n, bool swap, double slope, double intercept, int bits_stored, float* dst) {
    const bool identity = slope == 1.0 && intercept == 0.0;
    const int bits = (int)sizeof(S) * 8;
    const bool narrow = bits_stored > 0 && bits_stored < bits;
    for (size_t i = 0; i < n; ++i) {
        uint8_t b[sizeof(S)];
        memcpy(b, src + i * sizeof(S), sizeof(S));
        if (swap)
            for (size_t k = 0; k < sizeof(S) / 2; ++k) { const uint8_t t = b[k]; b[k] = b[sizeof(S) - 1 - k]; b[sizeof(S) - 1 - k] = t; }
        S v;
        memcpy(&v, b, sizeof(S));
        if (narrow) {
            using U = typename std::make_unsigned<S>::type;
            U u = static_cast<U>(v) & static_cast<U>((U(1) << bits_stored) - 1);
            if (std::is_signed<S>::value && (u >> (bits_stored - 1)) & 1) u |= static_cast<U>(~U(0) << bits_stored);
            v = static_cast<S>(u);
        }
        dst[i] = identity ? static_cast<float>(v) : static_cast<float>(static_cast<double>(v) * slope + intercept);
    }
}

// ---- encapsulated pixel data (PS3.5 A.4): items (FFFE,E000) -- the first is the Basic Offset Table, the rest are fragments
// of the one frame -- up to the sequence delimiter (FFFE,E0DD).  Returns the concatenated fragments.
int dcm_fragments(const uint8_t* p, size_t n, std::vector<uint8_t>& out, const char* path) {
    size_t pos = 0;
    bool first = true;
    out.clear();
    for (;;) {
        if (pos + 8 > n) return set_error(SVB_ERR_FORMAT, "%s: encapsulated pixel data ends without a sequence delimiter", path);
        const uint16_t g = (uint16_t)(p[pos] | (p[pos + 1] << 8)), e = (uint16_t)(p[pos + 2] | (p[pos + 3] << 8));
        const uint32_t len = (uint32_t)p[pos + 4] | (uint32_t)p[pos + 5] << 8 | (uint32_t)p[pos + 6] << 16 | (uint32_t)p[pos + 7] << 24;
        pos += 8;
        if (g == 0xFFFE && e == 0xE0DD) break;
        if (g != 0xFFFE || e != 0xE000) return set_error(SVB_ERR_FORMAT, "%s: unexpected tag (%04x,%04x) inside encapsulated pixel data", path, g, e);
        if (len == 0xFFFFFFFFu || pos + len > n) return set_error(SVB_ERR_FORMAT, "%s: pixel data fragment runs past the end of the file", path);
        if (!first) out.insert(out.end(), p + pos, p + pos + len);
        first = false;
        pos += len;
    }
    if (out.empty()) return set_error(SVB_ERR_FORMAT, "%s: encapsulated pixel data has no fragment", path);
    return SVB_OK;
}

// ---- RLE Lossless (PS3.5 Annex G): 64-byte header (segment count + 15 offsets), every segment one PackBits-coded BYTE plane of
// the frame, most significant byte first.  Output: little-endian samples of `bytes_per_sample` bytes.
int rle_decode(const std::vector<uint8_t>& cs, size_t n_px, int bytes_per_sample, std::vector<uint8_t>& raw, const char* path) {
    if (cs.size() < 64) return set_error(SVB_ERR_FORMAT, "%s: RLE frame shorter than its header", path);
    auto rd32 = [&](size_t o) { return (uint32_t)cs[o] | (uint32_t)cs[o + 1] << 8 | (uint32_t)cs[o + 2] << 16 | (uint32_t)cs[o + 3] << 24; };
    const uint32_t nseg = rd32(0);
    if ((int)nseg != bytes_per_sample) return set_error(SVB_ERR_FORMAT, "%s: RLE frame has %u segments, %d expected", path, nseg, bytes_per_sample);
    raw.assign(n_px * (size_t)bytes_per_sample, 0);
    for (uint32_t sgi = 0; sgi < nseg; ++sgi) {
        const size_t lo = rd32(4 + 4 * sgi), hi = sgi + 1 < nseg ? rd32(8 + 4 * sgi) : cs.size();
        if (lo < 64 || lo > hi || hi > cs.size()) return set_error(SVB_ERR_FORMAT, "%s: RLE segment %u has bad offsets", path, sgi);
        const int byte_pos = bytes_per_sample - 1 - (int)sgi;  // segment 0 = most significant byte
        size_t o = 0, q = lo;
        while (o < n_px && q < hi) {
            const int8_t c = (int8_t)cs[q++];
            if (c >= 0) {  // literal run of c + 1 bytes
                const size_t run = (size_t)c + 1;
                if (q + run > hi) return set_error(SVB_ERR_FORMAT, "%s: RLE literal run leaves its segment", path);
                for (size_t k = 0; k < run && o < n_px; ++k, ++o) raw[o * bytes_per_sample + byte_pos] = cs[q + k];
                q += run;
            } else if (c != -128) {  // replicate the next byte 1 - c times
                if (q >= hi) return set_error(SVB_ERR_FORMAT, "%s: RLE replicate run leaves its segment", path);
                const size_t run = (size_t)(1 - (int)c);
                const uint8_t v = cs[q++];
                for (size_t k = 0; k < run && o < n_px; ++k, ++o) raw[o * bytes_per_sample + byte_pos] = v;
            }
        }
        if (o < n_px) return set_error(SVB_ERR_FORMAT, "%s: RLE segment %u decodes to %zu of %zu bytes", path, sgi, o, n_px);
    }
    return SVB_OK;
}

// ---- JPEG Lossless, process 14 (ITU-T T.81 Annex H): SOF3, one component, Huffman-coded differences to one of the seven
// predictors, point transform, restart intervals.  Output: little-endian uint16 samples (or bytes for precision <= 8 when
// `bytes_per_sample` is 1).
struct JpgHuff {
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    bool present = false;
};
struct JpgBits {
    const uint8_t* p;
    size_t n, pos;
    uint32_t acc = 0;
    int cnt = 0;
    bool hit_marker = false;
    int bit() {
        if (cnt == 0) {
            uint8_t b = 0;
            if (pos < n && !hit_marker) {
                b = p[pos++];
                if (b == 0xFF) {
                    if (pos < n && p[pos] == 0x00) ++pos;  // stuffed zero
                    else { hit_marker = true; --pos; b = 0; }  // a marker: feed zeros (the caller resynchronises at RSTn)
                }
            }
            acc = b;
            cnt = 8;
        }
        --cnt;
        return (acc >> cnt) & 1;
    }
    int bits(int k) { int v = 0; while (k-- > 0) v = (v << 1) | bit(); return v; }
};
int jpeg_lossless_decode(const std::vector<uint8_t>& cs, int rows, int cols, int bytes_per_sample, std::vector<uint8_t>& raw, const char* path) {
    const uint8_t* p = cs.data();
    const size_t n = cs.size();
    if (n < 4 || p[0] != 0xFF || p[1] != 0xD8) return set_error(SVB_ERR_FORMAT, "%s: JPEG stream has no SOI marker", path);
    JpgHuff huff[4];
    int precision = 0, restart = 0, comp_table = 0, predictor = 0, pt = 0, width = 0, height = 0;
    size_t pos = 2, scan = 0;
    while (pos + 4 <= n && scan == 0) {
        if (p[pos] != 0xFF) return set_error(SVB_ERR_FORMAT, "%s: JPEG marker expected at byte %zu", path, pos);
        const uint8_t m = p[pos + 1];
        if (m == 0xFF) { ++pos; continue; }  // fill byte
        const size_t seg = ((size_t)p[pos + 2] << 8) | p[pos + 3];
        if (seg < 2 || pos + 2 + seg > n) return set_error(SVB_ERR_FORMAT, "%s: JPEG segment %02x runs past the stream", path, m);
        const uint8_t* d = p + pos + 4;
        const size_t dl = seg - 2;
        if (m == 0xC3) {  // SOF3: lossless, Huffman
            if (dl < 6 + 3 || d[5] != 1) return set_error(SVB_ERR_FORMAT, "%s: lossless JPEG with %d components (1 supported)", path, dl >= 6 ? d[5] : 0);
            precision = d[0]; height = (d[1] << 8) | d[2]; width = (d[3] << 8) | d[4];
        } else if (m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return set_error(SVB_ERR_FORMAT, "%s: JPEG frame type SOF%d is not lossless Huffman (SOF3)", path, m - 0xC0);
        } else if (m == 0xC4) {  // DHT: one or more tables
            size_t q = 0;
            while (q + 17 <= dl) {
                const int tc = d[q] >> 4, th = d[q] & 15;
                if (tc != 0 || th > 3) return set_error(SVB_ERR_FORMAT, "%s: bad Huffman table id %02x", path, d[q]);
                int counts[17] = {0}, total = 0;
                for (int i = 1; i <= 16; ++i) { counts[i] = d[q + i]; total += counts[i]; }
                if (total > 256 || q + 17 + (size_t)total > dl) return set_error(SVB_ERR_FORMAT, "%s: Huffman table runs past its segment", path);
                JpgHuff& h = huff[th];
                memcpy(h.vals, d + q + 17, (size_t)total);
                int code = 0, k = 0;
                for (int i = 1; i <= 16; ++i) {  // T.81 Annex C / F.2.2.3
                    h.valptr[i] = k;
                    h.mincode[i] = code;
                    code += counts[i];
                    k += counts[i];
                    h.maxcode[i] = counts[i] ? code - 1 : -1;
                    code <<= 1;
                }
                h.maxcode[17] = 0x7FFFFFFF;
                h.present = true;
                q += 17 + (size_t)total;
            }
        } else if (m == 0xDD) {  // DRI
            if (dl >= 2) restart = (d[0] << 8) | d[1];
        } else if (m == 0xDA) {  // SOS
            if (dl < 6 || d[0] != 1) return set_error(SVB_ERR_FORMAT, "%s: scan with %d components (1 supported)", path, dl ? d[0] : 0);
            comp_table = d[2] >> 4;
            predictor = d[3];
            pt = d[5] & 15;
            scan = pos + 2 + seg;
        }
        pos += 2 + seg;
    }
    if (!scan || precision < 2 || precision > 16) return set_error(SVB_ERR_FORMAT, "%s: lossless JPEG without SOF3 / SOS (precision %d)", path, precision);
    if (width != cols || height != rows) return set_error(SVB_ERR_FORMAT, "%s: JPEG frame is %dx%d, the DICOM header says %dx%d", path, height, width, rows, cols);
    if (predictor < 1 || predictor > 7) return set_error(SVB_ERR_FORMAT, "%s: lossless predictor %d", path, predictor);
    if (comp_table > 3 || !huff[comp_table].present) return set_error(SVB_ERR_FORMAT, "%s: scan refers to a missing Huffman table", path);
    if (precision > 8 * bytes_per_sample) return set_error(SVB_ERR_FORMAT, "%s: JPEG precision %d exceeds BitsAllocated %d", path, precision, 8 * bytes_per_sample);
    if (restart && restart % cols != 0) return set_error(SVB_ERR_FORMAT, "%s: restart interval %d is not a whole number of lines", path, restart);
    const JpgHuff& h = huff[comp_table];
    std::vector<uint16_t> img((size_t)rows * cols);
    JpgBits br{p, n, scan};
    const int init = 1 << (precision - pt - 1);
    long long in_interval = 0;
    bool fresh = true;  // the next line starts a restart interval (or the scan): predict as on the first line
    for (int y = 0; y < rows; ++y) {
        if (restart && y > 0 && in_interval == restart) {  // expect RSTn, byte-align, reset the prediction
            br.cnt = 0;
            if (!br.hit_marker) {  // skip to the marker (padding bits were 1s inside the last byte)
                while (br.pos + 1 < n && !(p[br.pos] == 0xFF && p[br.pos + 1] >= 0xD0 && p[br.pos + 1] <= 0xD7)) ++br.pos;
            }
            if (br.pos + 1 >= n || p[br.pos] != 0xFF || p[br.pos + 1] < 0xD0 || p[br.pos + 1] > 0xD7)
                return set_error(SVB_ERR_FORMAT, "%s: restart marker missing before line %d", path, y);
            br.pos += 2;
            br.hit_marker = false;
            in_interval = 0;
            fresh = true;
        }
        uint16_t* row = img.data() + (size_t)y * cols;
        const uint16_t* up = y > 0 ? row - cols : nullptr;
        for (int x = 0; x < cols; ++x) {
            int code = 0, len = 0, s = -1;  // DECODE (T.81 F.2.2.3)
            for (len = 1; len <= 16; ++len) {
                code = (code << 1) | br.bit();
                if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) { s = h.vals[h.valptr[len] + code - h.mincode[len]]; break; }
            }
            if (s < 0 || s > 16) return set_error(SVB_ERR_FORMAT, "%s: corrupt Huffman code at line %d column %d", path, y, x);
            int diff;
            if (s == 0) diff = 0;
            else if (s == 16) diff = 32768;
            else {
                diff = br.bits(s);
                if (diff < (1 << (s - 1))) diff -= (1 << s) - 1;  // EXTEND
            }
            int pred;
            if (fresh) pred = x == 0 ? init : row[x - 1];            // first line of the scan / of a restart interval
            else if (x == 0) pred = up[0];                            // first sample of the other lines: the one above
            else {
                const int ra = row[x - 1], rb = up[x], rc = up[x - 1];
                switch (predictor) {
                    case 1: pred = ra; break;
                    case 2: pred = rb; break;
                    case 3: pred = rc; break;
                    case 4: pred = ra + rb - rc; break;
                    case 5: pred = ra + ((rb - rc) >> 1); break;
                    case 6: pred = rb + ((ra - rc) >> 1); break;
                    default: pred = (ra + rb) >> 1; break;
                }
            }
            row[x] = (uint16_t)((pred + diff) & 0xFFFF);
        }
        if (br.hit_marker && !(restart && y + 1 < rows && in_interval + cols == restart) && y + 1 < rows)
            return set_error(SVB_ERR_FORMAT, "%s: JPEG entropy data ends at line %d of %d", path, y + 1, rows);
        in_interval += cols;
        fresh = false;
    }
    raw.resize((size_t)rows * cols * bytes_per_sample);
    for (size_t i = 0; i < (size_t)rows * cols; ++i) {
        const uint16_t v = (uint16_t)(img[i] << pt);
        raw[i * bytes_per_sample] = (uint8_t)v;
        if (bytes_per_sample > 1) raw[i * bytes_per_sample + 1] = (uint8_t)(v >> 8);
    }
    return SVB_OK;
}

int dicom_pixels(const char* path, const svb_dicom_info* info, float* dst, size_t dst_elems) {
    if (!path || !info || !dst) return set_error(SVB_ERR_INVALID_ARG, "dicom: null argument");
    if (info->rows <= 0 || info->cols <= 0) return set_error(SVB_ERR_INVALID_ARG, "dicom: bad size %d x %d", info->rows, info->cols);
    const size_t n = (size_t)info->rows * info->cols;
    if (dst_elems < n) return set_error(SVB_ERR_WORKSPACE_TOO_SMALL, "dicom: destination holds %zu elements, slice has %zu", dst_elems, n);
    const int bps = info->bits_allocated / 8;
    const size_t bytes = n * (size_t)bps;
    const size_t to_read = info->encapsulation != SVB_DICOM_NATIVE ? (size_t)(info->pixel_bytes > 0 ? info->pixel_bytes : 0) : bytes;
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(SVB_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    std::vector<uint8_t> raw(to_read);
    const bool ok = fseek(f, (long)info->pixel_offset, SEEK_SET) == 0 && fread(raw.data(), 1, to_read, f) == to_read;
    fclose(f);
    if (!ok) return set_error(SVB_ERR_IO, "%s: pixel data truncated", path);
    bool swap;
    if (info->encapsulation != SVB_DICOM_NATIVE) {
        std::vector<uint8_t> cs, dec;
        if (int rc = dcm_fragments(raw.data(), raw.size(), cs, path)) return rc;
        if (info->encapsulation == SVB_DICOM_RLE) {
            if (int rc = rle_decode(cs, n, bps, dec, path)) return rc;
        } else {
            if (int rc = jpeg_lossless_decode(cs, info->rows, info->cols, bps, dec, path)) return rc;
        }
        raw.swap(dec);  // little-endian samples
        const uint16_t probe = 1;
        swap = *reinterpret_cast<const uint8_t*>(&probe) == 0;
    } else {
        const uint16_t probe = 1;
        const bool host_big = *reinterpret_cast<const uint8_t*>(&probe) == 0;
        swap = (info->big_endian != 0) != host_big;
    }
    const double sl = info->rescale_slope, ic = info->rescale_intercept;
    const bool sgn = info->pixel_representation != 0;
    const int bs = info->bits_stored;
    switch (info->bits_allocated) {
        case 8: sgn ? convert_rescaled<int8_t>(raw.data(), n, false, sl, ic, bs, dst) : convert_rescaled<uint8_t>(raw.data(), n, false, sl, ic, bs, dst); break;
        case 16: sgn ? convert_rescaled<int16_t>(raw.data(), n, swap, sl, ic, bs, dst) : convert_rescaled<uint16_t>(raw.data(), n, swap, sl, ic, bs, dst); break;
        default: sgn ? convert_rescaled<int32_t>(raw.data(), n, swap, sl, ic, bs, dst) : convert_rescaled<uint32_t>(raw.data(), n, swap, sl, ic, bs, dst); break;
    }
    return SVB_OK;
}

}  // namespace

extern "C" {

size_t svb_png_bound(int h, int w) {
    if (h <= 0 || w <= 0) return 0;
    const size_t raw = (size_t)h * ((size_t)w + 1);
    return 8 + 25 + 12 + (size_t)compressBound((uLong)raw) + 12;
}

int svb_png_encode_gray8(const uint8_t* h_img, int h, int w, int level, uint8_t* h_out, size_t cap, size_t* out_len) {
    return guarded("png encode", [&] { return png_encode(h_img, h, w, level, h_out, cap, out_len); });
}

int svb_png_write_gray8_batch(const uint8_t* h_imgs, int n, int h, int w, const char* const* paths, int level, int n_threads,
                              int32_t* rcs) {
    if (n < 0 || (n > 0 && (!h_imgs || !paths)) || h <= 0 || w <= 0)
        return set_error(SVB_ERR_INVALID_ARG, "png batch: bad arguments (n=%d h=%d w=%d)", n, h, w);
    const size_t cap = svb_png_bound(h, w);
    std::atomic<int> first_bad{-1};
    std::vector<std::string> msgs((size_t)(n > 0 ? n : 0));
    parallel_for(n, n_threads, [&](int i) {
        int rc = guarded("png batch", [&] {
            std::vector<uint8_t> buf(cap);
            size_t len = 0;
            int r = png_encode(h_imgs + (size_t)i * h * w, h, w, level, buf.data(), cap, &len);
            if (r == SVB_OK) r = write_file(paths[i], buf.data(), len);
            return r;
        });
        if (rcs) rcs[i] = rc;
        if (rc != SVB_OK) {
            msgs[i] = svb_last_error();  // thread-local in the worker
            int expect = -1;
            first_bad.compare_exchange_strong(expect, i);
        }
    });
    const int bad = first_bad.load();
    if (bad >= 0) return set_error(SVB_ERR_IO, "png batch: image %d failed: %s", bad, msgs[bad].c_str());
    return SVB_OK;
}

int svb_png_write_gray8_ragged(const uint8_t* h_pool, const int64_t* offs, const int32_t* hw, int n, const char* const* paths,
                               int level, int n_threads, int32_t* rcs) {
    if (n < 0 || (n > 0 && (!h_pool || !offs || !hw || !paths)))
        return set_error(SVB_ERR_INVALID_ARG, "png ragged batch: bad arguments (n=%d)", n);
    std::atomic<int> first_bad{-1};
    std::vector<std::string> msgs((size_t)(n > 0 ? n : 0));
    parallel_for(n, n_threads, [&](int i) {
        const int h = hw[2 * i], w = hw[2 * i + 1];
        int rc = SVB_OK;
        if (h <= 0 || w <= 0) rc = set_error(SVB_ERR_INVALID_ARG, "png ragged batch: image %d has size %d x %d", i, h, w);
        if (rc == SVB_OK)
            rc = guarded("png ragged batch", [&] {
                const size_t cap = svb_png_bound(h, w);
                std::vector<uint8_t> buf(cap);
                size_t len = 0;
                int r = png_encode(h_pool + offs[i], h, w, level, buf.data(), cap, &len);
                if (r == SVB_OK) r = write_file(paths[i], buf.data(), len);
                return r;
            });
        if (rcs) rcs[i] = rc;
        if (rc != SVB_OK) {
            msgs[i] = svb_last_error();
            int expect = -1;
            first_bad.compare_exchange_strong(expect, i);
        }
    });
    const int bad = first_bad.load();
    if (bad >= 0) return set_error(SVB_ERR_IO, "png ragged batch: image %d failed: %s", bad, msgs[bad].c_str());
    return SVB_OK;
}

int svb_mha_read_header(const char* path, svb_mha_info* info) {
    return guarded("mha header", [&] { return mha_header(path, info); });
}

int svb_mha_read_f32(const char* path, const svb_mha_info* info, float* h_dst, size_t dst_elems) {
    return guarded("mha read", [&] { return mha_read(path, info, h_dst, dst_elems); });
}

int svb_mha_read_batch_f32(const char* const* paths, int n, const svb_mha_info* infos, float* const* h_dsts,
                           const size_t* dst_elems, int n_threads, int32_t* rcs) {
    if (n < 0 || (n > 0 && (!paths || !infos || !h_dsts || !dst_elems)))
        return set_error(SVB_ERR_INVALID_ARG, "mha batch: bad arguments (n=%d)", n);
    std::atomic<int> n_bad{0};
    parallel_for(n, n_threads, [&](int i) {
        const int rc = guarded("mha batch", [&] { return mha_read(paths[i], &infos[i], h_dsts[i], dst_elems[i]); });
        if (rcs) rcs[i] = rc;
        if (rc != SVB_OK) n_bad.fetch_add(1);
    });
    // per-file failures are the caller's to skip (the reference's drivers skip unreadable series: spider.py:139-141)
    return n_bad.load() ? set_error(SVB_ERR_IO, "mha batch: %d of %d volumes failed (see the per-file codes)", n_bad.load(), n) : SVB_OK;
}

int svb_mha_read_batch_slab_f32(const char* const* paths, int n, const svb_mha_info* infos, float* const* h_dsts,
                                const size_t* dst_elems, const int32_t* z0, const int32_t* z1, int n_threads, int32_t* rcs) {
    if (n < 0 || (n > 0 && (!paths || !infos || !h_dsts || !dst_elems || !z0 || !z1)))
        return set_error(SVB_ERR_INVALID_ARG, "mha slab batch: bad arguments (n=%d)", n);
    std::atomic<int> n_bad{0};
    parallel_for(n, n_threads, [&](int i) {
        const int rc = guarded("mha slab batch", [&] { return mha_read_slab(paths[i], &infos[i], h_dsts[i], dst_elems[i], z0[i], z1[i]); });
        if (rcs) rcs[i] = rc;
        if (rc != SVB_OK) n_bad.fetch_add(1);
    });
    return n_bad.load() ? set_error(SVB_ERR_IO, "mha slab batch: %d of %d volumes failed (see the per-file codes)", n_bad.load(), n) : SVB_OK;
}

int svb_dicom_read_headers(const char* const* paths, int n, svb_dicom_info* infos, int n_threads, int32_t* rcs) {
    if (n < 0 || (n > 0 && (!paths || !infos))) return set_error(SVB_ERR_INVALID_ARG, "dicom headers: bad arguments (n=%d)", n);
    std::atomic<int> n_bad{0}, first_bad{-1};
    std::vector<std::string> msgs((size_t)(n > 0 ? n : 0));
    parallel_for(n, n_threads, [&](int i) {
        const int rc = guarded("dicom headers", [&] { return dicom_header(paths[i], &infos[i], nullptr); });
        if (rcs) rcs[i] = rc;
        if (rc != SVB_OK) {
            n_bad.fetch_add(1);
            msgs[i] = svb_last_error();  // thread-local in the worker
            int expect = -1;
            first_bad.compare_exchange_strong(expect, i);
        }
    });
    // files that are not DICOM (or not a supported one) are the caller's to drop, like GDCM's directory scan does
    return n_bad.load() ? set_error(SVB_ERR_FORMAT, "dicom headers: %d of %d files are not supported DICOM slices; first: %s", n_bad.load(), n,
                                    msgs[first_bad.load()].c_str())
                        : SVB_OK;
}

int svb_dicom_read_slices_f32(const char* const* paths, int n, const svb_dicom_info* infos, float* const* h_dsts,
                              const size_t* dst_elems, int n_threads, int32_t* rcs) {
    if (n < 0 || (n > 0 && (!paths || !infos || !h_dsts || !dst_elems)))
        return set_error(SVB_ERR_INVALID_ARG, "dicom slices: bad arguments (n=%d)", n);
    std::atomic<int> first_bad{-1};
    std::vector<std::string> msgs((size_t)(n > 0 ? n : 0));
    parallel_for(n, n_threads, [&](int i) {
        const int rc = guarded("dicom slices", [&] { return dicom_pixels(paths[i], &infos[i], h_dsts[i], dst_elems[i]); });
        if (rcs) rcs[i] = rc;
        if (rc != SVB_OK) {
            msgs[i] = svb_last_error();
            int expect = -1;
            first_bad.compare_exchange_strong(expect, i);
        }
    });
    const int bad = first_bad.load();
    if (bad >= 0) return set_error(SVB_ERR_IO, "dicom slices: file %d failed: %s", bad, msgs[bad].c_str());
    return SVB_OK;
}

}  // extern "C"
