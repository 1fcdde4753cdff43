#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <chrono>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
using clk = std::chrono::steady_clock;
int main(int argc, char **argv) {
    const char *path = argv[1];
    int nt = argc > 2 ? atoi(argv[2]) : 16;
    int fd = open(path, O_RDONLY);
    struct stat st; fstat(fd, &st);
    size_t n = st.st_size;
    char *dst = (char *)aligned_alloc(4096, n);
    memset(dst, 1, n);   // touch
    for (int mode = 0; mode < 4; ++mode) {
        const char *src = nullptr;
        if (mode == 0 || mode == 2 || mode == 3) {
            src = (const char *)mmap(nullptr, n, PROT_READ, MAP_SHARED | (mode == 2 ? MAP_POPULATE : 0), fd, 0);
            if (mode == 3) madvise((void *)src, n, MADV_HUGEPAGE);
        }
        auto t0 = clk::now();
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) pool.emplace_back([&, t]() {
            size_t a = n * t / nt, b = n * (t + 1) / nt;
            if (mode == 1) { size_t off = a; while (off < b) { ssize_t r = pread(fd, dst + off, std::min<size_t>(b - off, 8 << 20), off); if (r <= 0) break; off += r; } }
            else memcpy(dst + a, src + a, b - a);
        });
        for (auto &th : pool) th.join();
        double dt = std::chrono::duration<double>(clk::now() - t0).count();
        printf("mode %d (%s) %d threads: %.2f GB in %.3f s = %.1f GB/s\n", mode,
               mode == 0 ? "mmap memcpy" : mode == 1 ? "pread" : mode == 2 ? "mmap populate (map time excluded)" : "mmap hugepage advice", nt, n / 1e9, dt, n / 1e9 / dt);
        if (src) munmap((void *)src, n);
    }
}
