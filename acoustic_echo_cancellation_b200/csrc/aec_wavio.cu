// Batched wav ingest of the h5 generators (host code, no CUDA): RIFF/WAVE probe and a multi-threaded reader that
// puts 16-bit PCM samples straight into the rows of a (page-locked) int16 batch buffer.
// Replaces the four serial `librosa.load` calls per utterance of the reference's generators
// (Stage2_lhm/generate_h5files/train_wav2h5.py:20-23, test_wav2h5.py:29-32, val_wav2h5.py:33-36) for the corpus
// format they are run on (16-bit PCM, mono, already at --sr); anything else is reported as AEC_EUNSUPPORTED and the
// Python host side falls back to the general loader.  No sample arithmetic happens here: int16 in, int16 out.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/aec_b200.h"

namespace {

inline uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

int probe_fd(int fd, aec_wav_info* info) {
    struct stat st;
    if (fstat(fd, &st) != 0) return AEC_EIO;
    unsigned char head[12];
    if (pread(fd, head, 12, 0) != 12 || memcmp(head, "RIFF", 4) != 0 || memcmp(head + 8, "WAVE", 4) != 0) return AEC_EINVAL;
    int64_t off = 12;
    bool have_fmt = false;
    memset(info, 0, sizeof(*info));
    for (;;) {
        unsigned char ch[8];
        if (pread(fd, ch, 8, off) != 8) return AEC_EINVAL;          // no data chunk
        const uint32_t size = rd32(ch + 4);
        off += 8;
        if (memcmp(ch, "fmt ", 4) == 0) {
            unsigned char body[40];
            const size_t want = size < sizeof(body) ? size : sizeof(body);
            if (size < 16 || pread(fd, body, want, off) != (ssize_t)want) return AEC_EINVAL;
            info->format = rd16(body);
            info->channels = rd16(body + 2);
            info->rate = (int32_t)rd32(body + 4);
            info->bits = rd16(body + 14);
            if (info->format == 0xFFFE && size >= 26) info->format = rd16(body + 24);   // WAVE_FORMAT_EXTENSIBLE
            have_fmt = true;
        } else if (memcmp(ch, "data", 4) == 0) {
            if (!have_fmt) return AEC_EINVAL;
            int64_t bytes = size;
            const int64_t avail = (int64_t)st.st_size - off;
            if (bytes > avail) bytes = avail;                        // (streamed files carry 0xFFFFFFFF here)
            const int64_t fb = (int64_t)info->channels * (info->bits / 8);
            info->frames = fb > 0 ? bytes / fb : 0;
            info->data_offset = off;
            return AEC_OK;
        }
        off += (int64_t)size + (size & 1);
    }
}

template <typename F>
void parallel_for(int64_t n, int threads, F&& body) {
    if (threads < 1) threads = 1;
    if (threads > n) threads = (int)n;
    if (threads <= 1) {
        for (int64_t i = 0; i < n; ++i) body(i);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (int64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) body(i);
        });
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int aec_wav_probe(const char* path, aec_wav_info* info) {
    if (!path || !info) return AEC_EINVAL;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return AEC_EIO;
    const int rc = probe_fd(fd, info);
    close(fd);
    return rc;
}

extern "C" int aec_wav_probe_batch(const char* const* paths, int64_t n, aec_wav_info* infos, int32_t threads) {
    if (n < 0 || (n > 0 && (!paths || !infos))) return AEC_EINVAL;
    std::atomic<int> first_err{AEC_OK};
    parallel_for(n, threads, [&](int64_t i) {
        const int rc = aec_wav_probe(paths[i], &infos[i]);
        int ok = AEC_OK;
        if (rc != AEC_OK) first_err.compare_exchange_strong(ok, rc);
    });
    return first_err.load();
}

extern "C" int aec_wav_read_pcm16_batch(const char* const* paths, int64_t n, int16_t* dst, int64_t row_stride,
                                        int64_t row_samples, int64_t* frames, int32_t expect_rate, int32_t threads) {
    if (n < 0 || row_samples < 0 || row_stride < row_samples || (n > 0 && (!paths || !dst))) return AEC_EINVAL;
    std::atomic<int> first_err{AEC_OK};
    parallel_for(n, threads, [&](int64_t i) {
        auto fail = [&](int rc) {
            int ok = AEC_OK;
            first_err.compare_exchange_strong(ok, rc);
        };
        const int fd = open(paths[i], O_RDONLY);
        if (fd < 0) return fail(AEC_EIO);
        aec_wav_info info;
        int rc = probe_fd(fd, &info);
        if (rc == AEC_OK && !(info.format == 1 && info.bits == 16 && info.channels == 1 &&
                              (expect_rate <= 0 || info.rate == expect_rate)))
            rc = AEC_EUNSUPPORTED;
        if (rc != AEC_OK) {
            close(fd);
            return fail(rc);
        }
        if (frames) frames[i] = info.frames;
        int16_t* row = dst + i * row_stride;
        const int64_t take = info.frames < row_samples ? info.frames : row_samples;
        int64_t got = 0;
        while (got < take * 2) {
            const ssize_t k = pread(fd, reinterpret_cast<char*>(row) + got, (size_t)(take * 2 - got), info.data_offset + got);
            if (k <= 0) {
                close(fd);
                return fail(AEC_EIO);
            }
            got += k;
        }
        close(fd);
        if (take < row_samples) memset(row + take, 0, (size_t)(row_samples - take) * sizeof(int16_t));
    });
    return first_err.load();
}
