// Per-device pools of device and pinned-host blocks.
//
// The reference builds one Scene and renders it once per WorldEnd (src/core/api.rs:1715-1755); a host that renders
// scene after scene through this library would otherwise pay cudaMalloc/cudaFree (and cudaMallocHost) for ~6 GB of
// path state, the scene tables and the film on every call -- each of those calls synchronises the device and costs
// milliseconds.  Blocks released by pbrt_b200_scene_destroy stay in the pool and are handed to the next scene.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace pb {

// Device block of at least `bytes` (256-byte aligned) on the CURRENT device.  Returns nullptr on failure (cudaGetLastError
// carries the reason).  *got = real size of the block (pass it back to pool_free).
void* pool_alloc(size_t bytes, size_t* got);
void pool_free(void* p, size_t bytes);
// Pinned host memory.
void* pool_alloc_host(size_t bytes, size_t* got);
void pool_free_host(void* p, size_t bytes);
// cudaFree / cudaFreeHost everything cached (all devices).
void pool_trim();

// Bump sub-allocator over one pooled block.
struct Arena {
    char* base = nullptr;
    size_t size = 0, used = 0;
    template <typename T> T* take(size_t count) {
        size_t off = (used + 255) & ~(size_t)255;
        size_t bytes = count * sizeof(T);
        if (off + bytes > size) return nullptr;
        used = off + bytes;
        return reinterpret_cast<T*>(base + off);
    }
    static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
};

}  // namespace pb
