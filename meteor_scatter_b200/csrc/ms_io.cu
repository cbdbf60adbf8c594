// A-io on the fast path: read the sample payload of many WAV files straight into caller-provided
// (pinned) rows with a pool of native threads -- one pread() loop per file, no interpreter lock, no intermediate
// buffer.  Stands where the reference calls scipy.io.wavfile.read once per file (dsp/src/main.py:249); the RIFF
// headers are parsed by the host binding (wavio.wav_info), this entry point only moves bytes.
#include <fcntl.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "ms_common.cuh"

extern "C" int ms_read_files(const char* const* paths, const int64_t* offsets, const int64_t* n_bytes, void* const* dst,
                             const int64_t* dst_capacity, int32_t n_files, int32_t n_threads) {
    MS_REQUIRE(n_files >= 0 && (n_files == 0 || (paths && offsets && n_bytes && dst && dst_capacity)), MS_ERR_INVALID_ARG,
               "ms_read_files: null pointer");
    if (n_files == 0) return MS_OK;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_files) n_threads = n_files;
    std::atomic<int> next(0), failed(-1);
    std::vector<std::string> errors((size_t)n_threads);
    auto work = [&](int tid) {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n_files || failed.load() >= 0) return;
            const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
            if (fd < 0) {
                errors[tid] = std::string("cannot open ") + paths[i] + ": " + strerror(errno);
                failed.store(i);
                return;
            }
            char* out = static_cast<char*>(dst[i]);
            int64_t got = 0;
            const int64_t want = n_bytes[i];
            while (got < want) {
                const ssize_t k = pread(fd, out + got, (size_t)(want - got), (off_t)(offsets[i] + got));
                if (k <= 0) {
                    errors[tid] = std::string(paths[i]) + ": file shorter than its data chunk";
                    failed.store(i);
                    break;
                }
                got += k;
            }
            close(fd);
            if (got == want && dst_capacity[i] > want) memset(out + want, 0, (size_t)(dst_capacity[i] - want));   // ragged tail
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    if (failed.load() >= 0) {
        for (const auto& e : errors)
            if (!e.empty()) {
                ms::set_error("ms_read_files: %s", e.c_str());
                break;
            }
        return MS_ERR_INVALID_ARG;
    }
    return MS_OK;
}
