// Batched writer of the per-utterance training files (host code, no CUDA): what replaces the
//     writer = h5py.File(tr_filename, 'w'); writer.create_dataset(key, data=x.astype(np.float32), shape=x.shape, chunks=True) x 4;
//     writer.close()
// block of the reference's generator (Stage2_lhm/generate_h5files/train_wav2h5.py:35-44) for a whole batch of utterances:
// one HDF5 file per utterance with a flat root group of 1-D float32 datasets, written by C++ threads (no interpreter, no
// libhdf5).  The byte layout is the one of acoustic_echo_cancellation_b200/h5lite.py (superblock v0, object header v1, symbol
// table group: B-tree v1 node + symbol-table node + local heap, contiguous data-layout v3) -- the two writers are kept
// byte-identical by tests/test_h5lite.py, and h5lite.py documents the format sections each structure follows.
// 16-bit PCM sources (what the wav decoder leaves in the batch buffers) are stored as float32 = sample / 32768, the
// scaling of `librosa.load` (train_wav2h5.py:20-23); float32 sources are stored as they are.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/uio.h>
#include <unistd.h>

#include "../../include/aec_b200.h"

namespace {

constexpr uint64_t kUndef = ~0ull;
constexpr int kLeafK = 4, kInternalK = 16;
constexpr size_t kSnodSize = 8 + 2 * kLeafK * 40;                              // 328
constexpr size_t kTreeSize = 24 + (2 * kInternalK + 1) * 8 + 2 * kInternalK * 8;   // 544

struct Buf {
    std::vector<unsigned char> b;
    void u8(unsigned v) { b.push_back((unsigned char)v); }
    void u16(unsigned v) { u8(v & 255); u8((v >> 8) & 255); }
    void u32(uint32_t v) { for (int i = 0; i < 4; ++i) u8((v >> (8 * i)) & 255); }
    void u64(uint64_t v) { for (int i = 0; i < 8; ++i) u8((unsigned)((v >> (8 * i)) & 255)); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void raw(const void* p, size_t n) { const unsigned char* c = (const unsigned char*)p; b.insert(b.end(), c, c + n); }
    void pad8() { zeros((8 - b.size() % 8) % 8); }
    size_t size() const { return b.size(); }
};

inline uint64_t pad8(uint64_t n) { return (n + 7) & ~7ull; }

// object header (version 1) of a 1-D float32 dataset of n elements stored contiguously at addr: 120 bytes
void dataset_header(Buf& o, uint64_t n, uint64_t addr) {
    o.u8(1); o.u8(0); o.u16(4); o.u32(1); o.u32(104); o.zeros(4);              // prefix: 4 messages, 104 bytes
    o.u16(0x0001); o.u16(16); o.u8(0); o.zeros(3);                              // dataspace v1, rank 1
    o.u8(1); o.u8(1); o.u8(0); o.u8(0); o.zeros(4); o.u64(n);
    o.u16(0x0003); o.u16(24); o.u8(1); o.zeros(3);                              // datatype v1: IEEE float32, little endian
    o.u8(0x11); o.u8(0x20); o.u8(31); o.u8(0); o.u32(4);
    o.u16(0); o.u16(32); o.u8(23); o.u8(8); o.u8(0); o.u8(23); o.u32(127); o.zeros(4);
    o.u16(0x0005); o.u16(8); o.u8(1); o.zeros(3);                               // fill value v2: default
    o.u8(2); o.u8(1); o.u8(2); o.u8(1); o.u32(0);
    o.u16(0x0008); o.u16(24); o.u8(0); o.zeros(3);                              // data layout v3: contiguous
    o.u8(3); o.u8(1); o.u64(n ? addr : kUndef); o.u64(n * 4); o.zeros(6);
}

bool write_all(int fd, const void* p, size_t n) {
    const char* c = (const char*)p;
    while (n) {
        const ssize_t k = write(fd, c, n);
        if (k <= 0) return false;
        c += k;
        n -= (size_t)k;
    }
    return true;
}

// one file; order[] = dataset indices in strcmp order of their names
int write_one(const char* path, int nd, const char* const* names, const int* order, const void* const* data,
              const int64_t* lens, const int32_t* formats, std::vector<float>& scratch) {
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return AEC_EIO;
    bool ok = true;
    static const unsigned char zero8[8] = {0};
    unsigned char sb[96];
    memset(sb, 0, sizeof(sb));
    ok = write_all(fd, sb, 96);                                                // superblock goes out last
    uint64_t pos = 96;
    std::vector<uint64_t> addr((size_t)nd, kUndef);
    for (int d = 0; d < nd && ok; ++d) {                                       // raw data, creation order, 8-aligned
        const int64_t n = lens[d];
        if (n <= 0) continue;
        const uint64_t a = pad8(pos);
        if (a != pos) ok = write_all(fd, zero8, (size_t)(a - pos));
        const float* src;
        if (formats && formats[d] == 1) {                                      // 16-bit PCM -> float32 = s / 32768
            scratch.resize((size_t)n);
            const int16_t* s = (const int16_t*)data[d];
            float* q = scratch.data();
            for (int64_t i = 0; i < n; ++i) q[i] = (float)s[i] * (1.0f / 32768.0f);
            src = q;
        } else {
            src = (const float*)data[d];
        }
        ok = ok && write_all(fd, src, (size_t)n * 4);
        addr[(size_t)d] = a;
        pos = a + (uint64_t)n * 4;
    }
    // metadata block
    const uint64_t base = pad8(pos);
    Buf m;
    std::vector<uint64_t> oh((size_t)nd), name_off((size_t)nd);
    Buf heap;
    heap.zeros(8);                                                             // offset 0: the empty name
    for (int i = 0; i < nd; ++i) {
        const int d = order[i];
        name_off[(size_t)i] = heap.size();
        heap.raw(names[d], strlen(names[d]) + 1);
        heap.pad8();
        oh[(size_t)i] = base + m.size();
        dataset_header(m, (uint64_t)(lens[d] > 0 ? lens[d] : 0), addr[(size_t)d]);
    }
    const uint64_t heap_addr = base + m.size();
    m.raw("HEAP", 4); m.u8(0); m.zeros(3); m.u64(heap.size()); m.u64(1); m.u64(heap_addr + 32);
    m.raw(heap.b.data(), heap.size());
    uint64_t snod_addr = 0;
    if (nd > 0) {
        snod_addr = base + m.size();
        const size_t start = m.size();
        m.raw("SNOD", 4); m.u8(1); m.u8(0); m.u16((unsigned)nd);
        for (int i = 0; i < nd; ++i) {
            m.u64(name_off[(size_t)i]); m.u64(oh[(size_t)i]); m.u32(0); m.u32(0); m.zeros(16);
        }
        m.zeros(kSnodSize - (m.size() - start));
    }
    const uint64_t tree_addr = base + m.size();
    {
        const size_t start = m.size();
        m.raw("TREE", 4); m.u8(0); m.u8(0); m.u16(nd > 0 ? 1 : 0); m.u64(kUndef); m.u64(kUndef); m.u64(0);
        if (nd > 0) { m.u64(snod_addr); m.u64(name_off[(size_t)nd - 1]); }
        m.zeros(kTreeSize - (m.size() - start));
    }
    const uint64_t root_oh = base + m.size();
    m.u8(1); m.u8(0); m.u16(1); m.u32(1); m.u32(24); m.zeros(4);
    m.u16(0x0011); m.u16(16); m.u8(0); m.zeros(3); m.u64(tree_addr); m.u64(heap_addr);
    const uint64_t eof = base + m.size();
    if (ok && base != pos) ok = write_all(fd, zero8, (size_t)(base - pos));
    ok = ok && write_all(fd, m.b.data(), m.size());
    // superblock version 0
    Buf s;
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    s.raw(sig, 8);
    s.u8(0); s.u8(0); s.u8(0); s.u8(0); s.u8(0); s.u8(8); s.u8(8); s.u8(0);
    s.u16(kLeafK); s.u16(kInternalK); s.u32(0);
    s.u64(0); s.u64(kUndef); s.u64(eof); s.u64(kUndef);
    s.u64(0); s.u64(root_oh); s.u32(1); s.u32(0); s.u64(tree_addr); s.u64(heap_addr);
    ok = ok && s.size() == 96 && pwrite(fd, s.b.data(), 96, 0) == 96;
    if (close(fd) != 0) ok = false;
    return ok ? AEC_OK : AEC_EIO;
}

}  // namespace

extern "C" int aec_ex_write_batch(const char* const* paths, int64_t n_files, int32_t n_datasets, const char* const* names,
                                  const void* const* data, const int64_t* lens, const int32_t* formats, int32_t threads) {
    if (n_files < 0 || n_datasets < 0 || n_datasets > 2 * kLeafK) return AEC_EINVAL;   // one symbol-table node
    if (n_files > 0 && (!paths || (n_datasets > 0 && (!names || !data || !lens)))) return AEC_EINVAL;
    std::vector<int> order((size_t)n_datasets);
    for (int i = 0; i < n_datasets; ++i) {
        if (!names[i] || !names[i][0] || strchr(names[i], '/')) return AEC_EINVAL;
        if (formats && formats[i] != 0 && formats[i] != 1) return AEC_EINVAL;
        order[(size_t)i] = i;
    }
    std::sort(order.begin(), order.end(), [&](int a, int b) { return strcmp(names[a], names[b]) < 0; });
    for (int i = 1; i < n_datasets; ++i)
        if (strcmp(names[order[(size_t)i - 1]], names[order[(size_t)i]]) == 0) return AEC_EINVAL;   // duplicate name
    for (int64_t f = 0; f < n_files; ++f) {
        if (!paths[f]) return AEC_EINVAL;
        for (int d = 0; d < n_datasets; ++d)
            if (lens[f * n_datasets + d] > 0 && !data[f * n_datasets + d]) return AEC_EINVAL;
    }
    if (threads < 1) threads = 1;
    if (threads > n_files) threads = (int)(n_files > 0 ? n_files : 1);
    std::atomic<int64_t> next{0};
    std::atomic<int> first_err{AEC_OK};
    auto worker = [&] {
        std::vector<float> scratch;
        for (int64_t f = next.fetch_add(1); f < n_files; f = next.fetch_add(1)) {
            const int rc = write_one(paths[f], n_datasets, names, order.data(), data + f * n_datasets, lens + f * n_datasets,
                                     formats, scratch);
            int ok = AEC_OK;
            if (rc != AEC_OK) first_err.compare_exchange_strong(ok, rc);
        }
    };
    if (threads <= 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        pool.reserve((size_t)threads);
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    return first_err.load();
}
