#include "pool.h"

#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/pbrt_b200.h"

namespace pb {
namespace {

struct Block { void* p; size_t bytes; int device; };
struct Pool {
    std::mutex mu;
    std::vector<Block> free_dev, free_host;
    size_t cached_dev = 0, cached_host = 0;
    std::unordered_map<void*, int> live_dev;  // device of every block handed out: pool_free must not trust cudaGetDevice()
};
Pool& pool() { static Pool* p = new Pool(); return *p; }  // leaked on purpose: CUDA may already be torn down at exit

// Keep at most this much idle memory per process; anything beyond is returned to the driver right away.  The device cap
// (default 48 GB, PBRT_B200_POOL_CACHE_GB overrides; 0 = cache nothing) matters to whoever shares the process: torch / NCCL
// allocate outside this pool and cannot reclaim what sits idle here -- pbrt_b200_release_cached_memory() hands it back.
const size_t kMaxCachedHost = (size_t)2 << 30;
size_t max_cached_dev() {
    static const size_t v = [] {
        const char* e = getenv("PBRT_B200_POOL_CACHE_GB");
        double gb = e ? atof(e) : 48.0;
        if (!(gb >= 0.0)) gb = 48.0;
        return (size_t)(gb * (double)((size_t)1 << 30));
    }();
    return v;
}

// Smallest cached block with bytes <= size <= 2 * bytes (+ slack for small ones), same device.
int pick(std::vector<Block>& v, size_t bytes, int device) {
    int best = -1;
    for (int i = 0; i < (int)v.size(); ++i) {
        const Block& b = v[i];
        if (b.device != device || b.bytes < bytes || b.bytes > 2 * bytes + ((size_t)1 << 20)) continue;
        if (best < 0 || b.bytes < v[best].bytes) best = i;
    }
    return best;
}

}  // namespace

void* pool_alloc(size_t bytes, size_t* got) {
    if (bytes == 0) bytes = 256;
    bytes = (bytes + 255) & ~(size_t)255;
    int dev = 0;
    cudaGetDevice(&dev);
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> g(P.mu);
        int i = pick(P.free_dev, bytes, dev);
        if (i >= 0) {
            Block b = P.free_dev[i];
            P.free_dev.erase(P.free_dev.begin() + i);
            P.cached_dev -= b.bytes;
            P.live_dev[b.p] = dev;
            *got = b.bytes;
            return b.p;
        }
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        pool_trim();  // give cached blocks back and retry once
        if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    }
    {
        std::lock_guard<std::mutex> g(P.mu);
        P.live_dev[p] = dev;
    }
    *got = bytes;
    return p;
}

void pool_free(void* p, size_t bytes) {
    if (!p) return;
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> g(P.mu);
        int dev = -1;
        auto it = P.live_dev.find(p);
        if (it != P.live_dev.end()) { dev = it->second; P.live_dev.erase(it); }
        if (dev >= 0 && P.cached_dev + bytes <= max_cached_dev()) {  // a block this pool did not hand out is never cached
            P.free_dev.push_back(Block{p, bytes, dev});
            P.cached_dev += bytes;
            return;
        }
    }
    cudaFree(p);  // valid from any current device (unified addressing)
}

void* pool_alloc_host(size_t bytes, size_t* got) {
    if (bytes == 0) bytes = 256;
    bytes = (bytes + 4095) & ~(size_t)4095;
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> g(P.mu);
        int i = pick(P.free_host, bytes, -1);
        if (i >= 0) {
            Block b = P.free_host[i];
            P.free_host.erase(P.free_host.begin() + i);
            P.cached_host -= b.bytes;
            *got = b.bytes;
            return b.p;
        }
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    *got = bytes;
    return p;
}

void pool_free_host(void* p, size_t bytes) {
    if (!p) return;
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> g(P.mu);
        if (P.cached_host + bytes <= kMaxCachedHost) {
            P.free_host.push_back(Block{p, bytes, -1});
            P.cached_host += bytes;
            return;
        }
    }
    cudaFreeHost(p);
}

void pool_trim() {
    Pool& P = pool();
    std::vector<Block> dev, host;
    {
        std::lock_guard<std::mutex> g(P.mu);
        dev.swap(P.free_dev); host.swap(P.free_host);
        P.cached_dev = P.cached_host = 0;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (const Block& b : dev) { cudaSetDevice(b.device); cudaFree(b.p); }
    cudaSetDevice(cur);
    for (const Block& b : host) cudaFreeHost(b.p);
}

}  // namespace pb

extern "C" void pbrt_b200_release_cached_memory(void) { pb::pool_trim(); }
