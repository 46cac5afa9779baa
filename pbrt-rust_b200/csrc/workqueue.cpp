// Work counter shared between the processes that drive the GPUs of one box (include/pbrt_b200.h).
//
// The reference's SamplerIntegrator::render hands 16x16 tiles to its worker threads from one shared queue
// (src/core/integrator.rs:291-296: rayon's par_iter over the tile list).  Here the workers are one process per GPU, and
// the queue head is a 64-bit counter in POSIX shared memory (/dev/shm): a claim is ONE lock-free fetch-add on host memory,
// with no network round trip and no rank acting as a server.  Host-only code: it works without a GPU (the CPU tests drive
// it with the gloo backend).
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/pbrt_b200.h"
#include "error.h"

struct pbrt_b200_work_counter {
    std::atomic<uint64_t>* value = nullptr;  // in the shared mapping
    void* map = nullptr;
    std::string name;
};

namespace {
constexpr size_t kMapBytes = 4096;
static_assert(sizeof(std::atomic<uint64_t>) == 8 && std::atomic<uint64_t>::is_always_lock_free, "needs a lock-free 64-bit atomic in shared memory");
}  // namespace

extern "C" int pbrt_b200_work_counter_open(const char* name, int create, pbrt_b200_work_counter** out) {
    using pbrt_b200::fail;
    if (!name || !out || name[0] == '\0') return fail(PBRT_B200_ERR_INVALID, "work_counter_open: null argument");
    *out = nullptr;
    std::string n = name[0] == '/' ? std::string(name) : "/" + std::string(name);
    int fd = shm_open(n.c_str(), create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return fail(PBRT_B200_ERR_INVALID, "work_counter_open: shm_open(" + n + "): " + std::strerror(errno));
    if (create && ftruncate(fd, (off_t)kMapBytes) != 0) {
        int e = errno; close(fd); shm_unlink(n.c_str());
        return fail(PBRT_B200_ERR_INVALID, std::string("work_counter_open: ftruncate: ") + std::strerror(e));
    }
    void* p = mmap(nullptr, kMapBytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    int e = errno;
    close(fd);
    if (p == MAP_FAILED) return fail(PBRT_B200_ERR_INVALID, std::string("work_counter_open: mmap: ") + std::strerror(e));
    pbrt_b200_work_counter* c = new (std::nothrow) pbrt_b200_work_counter();
    if (!c) { munmap(p, kMapBytes); return fail(PBRT_B200_ERR_INVALID, "work_counter_open: out of memory"); }
    c->map = p; c->name = n;
    c->value = reinterpret_cast<std::atomic<uint64_t>*>(p);  // a fresh shm object is zero-filled: a valid atomic at 0
    if (create) c->value->store(0, std::memory_order_seq_cst);
    *out = c;
    return PBRT_B200_OK;
}

extern "C" uint64_t pbrt_b200_work_counter_fetch_add(pbrt_b200_work_counter* c, uint64_t n) {
    return c ? c->value->fetch_add(n, std::memory_order_acq_rel) : ~0ull;
}
extern "C" uint64_t pbrt_b200_work_counter_fetch_max(pbrt_b200_work_counter* c, uint64_t v) {
    if (!c) return ~0ull;
    uint64_t cur = c->value->load(std::memory_order_acquire);
    while (cur < v && !c->value->compare_exchange_weak(cur, v, std::memory_order_acq_rel, std::memory_order_acquire)) {}
    return cur;
}
extern "C" uint64_t pbrt_b200_work_counter_load(const pbrt_b200_work_counter* c) { return c ? c->value->load(std::memory_order_acquire) : ~0ull; }
extern "C" void pbrt_b200_work_counter_store(pbrt_b200_work_counter* c, uint64_t v) { if (c) c->value->store(v, std::memory_order_seq_cst); }
extern "C" void pbrt_b200_work_counter_close(pbrt_b200_work_counter* c, int unlink_name) {
    if (!c) return;
    munmap(c->map, kMapBytes);
    if (unlink_name) shm_unlink(c->name.c_str());
    delete c;
}
