// fmb_io.cu -- on-disk format of an index.
//
// The reference persists an index with cereal: BiFMIndex::serialize archives (bwt, C, annotatedArray, bwtRev)
// (fmindex/BiFMIndex.h:209-215), written and read by saveIndex / loadIndex (fmindex/diskStorage.h:13-27).  cereal is not
// available here and its archive layout depends on the String_c implementation, so the file written by this library holds
// the same CONTENT in a flat, versioned, layout-independent form: the BWT bytes of both directions and the sampled suffix
// array as "marker bitmap + samples in row order" -- exactly the arguments of fmb_index_create, i.e. of the reference's
// BiFMIndex(bwt, bwtRev, SparseArray) constructor (BiFMIndex.h:40-51).  C and every device table (occurrence blocks, pair
// table, k-mer tables, jump tables, locate tables) are rebuilt on the GPU at load time from those bytes, which takes
// seconds and keeps files small (8 GB instead of a 122 GB device image at 3 Gbp) and valid across layout changes.
//
//   offset  0  char[8]  magic "FMB200IX"
//           8  u32      version (2; version 1 = the same layout without header checksum, still read)
//          12  u32      sigma
//          16  u64      n (rows = text length incl. delimiters)
//          24  u32      kind: bit 0 = bwtRev present (BiFMIndex), bit 1 = no delimiter (NoDelim), bit 2 = ReuseRev (bidirectional, no bwtRev)
//          28  u32      header checksum: low 32 bits of fmb_checksum64 over the 120 header bytes with this field zero (version 1: 0)
//          32  u64      n_samples
//          40  u64[5]   section sizes in bytes: bwt, bwtRev, marker bitmap, sample seq ids, sample positions
//          80  u64[5]   section checksums (fmb_checksum64 below)
//         120  sections, back to back, in that order (little endian, u8 / u8 / u64 / u32 / u32 elements)
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "fmb_host.hpp"

using namespace fmb;

namespace {

constexpr char kMagic[8] = {'F', 'M', 'B', '2', '0', '0', 'I', 'X'};
constexpr uint32_t kVersion = 2;

struct FileHeader {
    char magic[8];
    uint32_t version, sigma;
    uint64_t n;
    uint32_t bidirectional, reserved;
    uint64_t n_samples;
    uint64_t bytes[5];
    uint64_t sum[5];
};
static_assert(sizeof(FileHeader) == 120, "header layout is part of the file format");

// order-sensitive 64-bit checksum over 8-byte little-endian words (the tail is zero padded): h = (h ^ w) * prime, FNV style
uint64_t checksum64(const void* data, uint64_t bytes) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint64_t h = 0xCBF29CE484222325ull;
    uint64_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x100000001B3ull;
    }
    if (i < bytes) {
        uint64_t w = 0;
        memcpy(&w, p + i, bytes - i);
        h = (h ^ w) * 0x100000001B3ull;
    }
    return h;
}

uint32_t header_checksum(FileHeader h) {
    h.reserved = 0;
    return (uint32_t)checksum64(&h, sizeof h);
}

struct FileCloser { void operator()(FILE* f) const { if (f) fclose(f); } };
using File = std::unique_ptr<FILE, FileCloser>;

}  // namespace

extern "C" {

uint64_t fmb_checksum64(const void* data, uint64_t bytes) { return checksum64(data, bytes); }

static int save_impl(const fmb_index* ix, const char* path);
static int load_impl(fmb_index** out, int device, const char* path);

// no C++ exception may cross the C boundary: the host buffers of a 3 Gbp index are 8 GB
int fmb_index_save(const fmb_index* ix, const char* path) {
    try {
        return save_impl(ix, path);
    } catch (const std::bad_alloc&) {
        set_error("fmb_index_save: out of host memory");
        return FMB_ENOMEM;
    } catch (const std::exception& e) {
        set_error("fmb_index_save: %s", e.what());
        return FMB_EINVAL;
    }
}
int fmb_index_load(fmb_index** out, int device, const char* path) {
    try {
        return load_impl(out, device, path);
    } catch (const std::bad_alloc&) {
        set_error("fmb_index_load: out of host memory");
        return FMB_ENOMEM;
    } catch (const std::exception& e) {
        set_error("fmb_index_load: %s", e.what());
        return FMB_EINVAL;
    }
}

static int save_impl(const fmb_index* ix, const char* path) {
    if (!ix || !path) { set_error("NULL argument"); return FMB_EINVAL; }
    const uint64_t n = ix->n, ns = ix->n_samples, words = (n + 63) / 64;
    std::vector<uint8_t> bwt(n), bwt_rev((ix->bidirectional && !ix->reuse_rev) ? n : 0);
    std::vector<uint64_t> bitmap(words);
    std::vector<uint32_t> seq(ns), pos(ns);
    FMB_TRY(fmb_index_export(ix, bwt.data(), (ix->bidirectional && !ix->reuse_rev) ? bwt_rev.data() : nullptr, bitmap.data(), seq.data(), pos.data()));
    FileHeader h{};
    memcpy(h.magic, kMagic, 8);
    h.version = kVersion;
    h.sigma = ix->sigma;
    h.n = n;
    const bool has_rev = ix->bidirectional && !ix->reuse_rev;
    h.bidirectional = (has_rev ? 1u : 0u) | (ix->first_symb ? 0u : 2u) | (ix->reuse_rev ? 4u : 0u);
    h.n_samples = ns;
    const void* sec[5] = {bwt.data(), bwt_rev.data(), bitmap.data(), seq.data(), pos.data()};
    h.bytes[0] = n; h.bytes[1] = bwt_rev.size(); h.bytes[2] = words * 8; h.bytes[3] = ns * 4; h.bytes[4] = ns * 4;
    for (int s = 0; s < 5; ++s) h.sum[s] = checksum64(sec[s], h.bytes[s]);
    h.reserved = header_checksum(h);
    // written next to the destination and renamed into place: a failed save never leaves a truncated file at `path`
    const std::string tmp = std::string(path) + ".tmp";
    File f(fopen(tmp.c_str(), "wb"));
    if (!f) { set_error("cannot open %s for writing", tmp.c_str()); return FMB_EINVAL; }
    bool ok = fwrite(&h, sizeof h, 1, f.get()) == 1;
    for (int s = 0; s < 5 && ok; ++s) ok = h.bytes[s] == 0 || fwrite(sec[s], 1, h.bytes[s], f.get()) == h.bytes[s];
    ok = ok && fflush(f.get()) == 0;
    ok = (fclose(f.release()) == 0) && ok;
    if (!ok) { remove(tmp.c_str()); set_error("short write to %s", tmp.c_str()); return FMB_EINVAL; }
    if (rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); set_error("cannot rename %s to %s", tmp.c_str(), path); return FMB_EINVAL; }
    return FMB_OK;
}

static int load_impl(fmb_index** out, int device, const char* path) {
    if (!out || !path) { set_error("NULL argument"); return FMB_EINVAL; }
    *out = nullptr;
    File f(fopen(path, "rb"));
    if (!f) { set_error("cannot open %s", path); return FMB_EINVAL; }
    FileHeader h{};
    if (fread(&h, sizeof h, 1, f.get()) != 1) { set_error("%s: truncated header", path); return FMB_EINVAL; }
    if (memcmp(h.magic, kMagic, 8) != 0) { set_error("%s: not an fmb200 index file (bad magic)", path); return FMB_EINVAL; }
    if (h.version != kVersion && h.version != 1) { set_error("%s: file format version %u, this library reads versions 1 and %u", path, h.version, kVersion); return FMB_EINVAL; }
    if (h.version >= 2 && h.reserved != header_checksum(h)) { set_error("%s: header checksum mismatch", path); return FMB_EINVAL; }
    if (h.version == 1 && h.reserved != 0) { set_error("%s: implausible header (reserved word set in a version 1 file)", path); return FMB_EINVAL; }
    // the range this build supports, checked before anything is allocated
    if (h.n >= 0xFFFFFFFFull - 64) { set_error("%s: n = %llu, this build supports n < 2^32 - 64", path, (unsigned long long)h.n); return FMB_EUNSUPPORTED; }
    const uint64_t words = (h.n + 63) / 64;
    const bool has_rev = (h.bidirectional & 1u) != 0;
    const uint64_t want[5] = {h.n, has_rev ? h.n : 0, words * 8, h.n_samples * 4, h.n_samples * 4};
    if (h.sigma < 2 || h.sigma > 32 || h.n == 0 || h.bidirectional > 7 || (has_rev && (h.bidirectional & 4u)) || (h.version == 1 && h.bidirectional > 1) ||
        h.n_samples > h.n) { set_error("%s: implausible header (sigma %u, n %llu)", path, h.sigma, (unsigned long long)h.n); return FMB_EINVAL; }
    for (int s = 0; s < 5; ++s)
        if (h.bytes[s] != want[s]) { set_error("%s: section %d holds %llu bytes, the header implies %llu", path, s, (unsigned long long)h.bytes[s], (unsigned long long)want[s]); return FMB_EINVAL; }
    {
        // the sizes the header implies must be the size of the file: nothing is allocated for a header that lies
        std::error_code ec;
        const uint64_t have = (uint64_t)std::filesystem::file_size(path, ec);
        uint64_t need = sizeof h;
        for (int s = 0; s < 5; ++s) need += want[s];
        if (!ec && have < need) { set_error("%s: truncated (%llu bytes, the header implies %llu)", path, (unsigned long long)have, (unsigned long long)need); return FMB_EINVAL; }
    }
    std::vector<uint8_t> bwt(h.n), bwt_rev(h.bytes[1]);
    std::vector<uint64_t> bitmap(words);
    std::vector<uint32_t> seq(h.n_samples), pos(h.n_samples);
    void* sec[5] = {bwt.data(), bwt_rev.data(), bitmap.data(), seq.data(), pos.data()};
    for (int s = 0; s < 5; ++s) {
        if (h.bytes[s] && fread(sec[s], 1, h.bytes[s], f.get()) != h.bytes[s]) { set_error("%s: truncated (section %d)", path, s); return FMB_EINVAL; }
        if (checksum64(sec[s], h.bytes[s]) != h.sum[s]) { set_error("%s: checksum mismatch in section %d", path, s); return FMB_EINVAL; }
    }
    if (fgetc(f.get()) != EOF) { set_error("%s: trailing bytes after the last section", path); return FMB_EINVAL; }
    return fmb_index_create_ex(out, device, h.sigma, h.n, bwt.data(), has_rev ? bwt_rev.data() : nullptr, bitmap.data(), seq.data(), pos.data(), h.n_samples,
                               ((h.bidirectional & 2u) ? FMB_INDEX_NO_DELIM : 0u) | ((h.bidirectional & 4u) ? FMB_INDEX_REUSE_REV : 0u));
}

}  // extern "C"
