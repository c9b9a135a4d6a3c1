/*
 * pool_emul.cpp — CPU test of the engine's device-memory pool (engine.cu MemPool) against a mock
 * CUDA allocator with a fixed capacity.  tests/test_native_abi.py extracts the struct's source text
 * from engine.cu into pool_extract.h (the product is not restructured for the test) and compiles this
 * file with it.  Checked: re-use by size class, the caps on cached big blocks, "the block released
 * last stays, older blocks of other sizes make room", trim-and-retry on allocation failure, and the
 * contract behind kept CUDA-IPC mappings: an exported block is never freed by release / trim /
 * eviction while pin_exported is set, only by trim_exported() and clear().
 * Test infrastructure only.
 */
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>
#include <vector>

/* ---- mock CUDA ---- */
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static size_t g_capacity = size_t(170) << 30, g_used = 0;
static std::map<void *, size_t> g_blocks;
static std::set<void *> g_pinned_by_peer; /* freed blocks a "peer" still maps: their memory is not returned */
static uintptr_t g_next = 0x10000000;
static long g_mallocs = 0, g_frees = 0;
static cudaError_t cudaMalloc(void **p, size_t bytes) {
    if (g_used + bytes > g_capacity) return 2;
    *p = reinterpret_cast<void *>(g_next);
    g_next += bytes + 4096;
    g_blocks[*p] = bytes;
    g_used += bytes;
    ++g_mallocs;
    return cudaSuccess;
}
static cudaError_t cudaFree(void *p) {
    auto it = g_blocks.find(p);
    if (it == g_blocks.end()) {
        std::printf("FAIL: cudaFree of an unknown pointer\n");
        std::exit(1);
    }
    if (!g_pinned_by_peer.count(p)) g_used -= it->second; /* a mapped block gives nothing back */
    g_blocks.erase(it);
    ++g_frees;
    return cudaSuccess;
}
static cudaError_t cudaGetLastError() { return cudaSuccess; }
enum { QGB_ERR_OOM = 5 };
static void fail(int, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw std::runtime_error(buf);
}

#include "pool_extract.h"

static int g_failures = 0;
#define CHECK(cond, ...)                      \
    do {                                      \
        if (!(cond)) {                        \
            ++g_failures;                     \
            std::printf("FAIL line %d: ", __LINE__); \
            std::printf(__VA_ARGS__);         \
            std::printf("\n");                \
        }                                     \
    } while (0)

static const size_t GiB = size_t(1) << 30;

int main() {
    { /* re-use by size class, small blocks are always cached */
        MemPool pool;
        void *a = pool.alloc(1000), *b = pool.alloc(600);
        CHECK(a != b && pool.live_bytes == 2048, "size classes: %zu live", pool.live_bytes);
        pool.release(a);
        void *c = pool.alloc(700);
        CHECK(c == a && g_mallocs == 2, "a released block of the class is re-used");
        pool.release(b);
        pool.release(c);
        pool.clear();
        CHECK(g_blocks.empty() && g_used == 0, "clear() frees everything");
    }
    { /* big blocks: two per size class, the rest is freed */
        MemPool pool;
        void *x[3] = {pool.alloc(16 * GiB), pool.alloc(16 * GiB), pool.alloc(16 * GiB)};
        for (void *p : x) pool.release(p);
        CHECK(pool.cached_bytes == 32 * GiB && g_used == 32 * GiB, "two 16 GiB blocks stay cached (%zu GiB)", pool.cached_bytes >> 30);
        /* the block released last stays: older blocks of other classes make room under the 136 GiB cap */
        void *big = pool.alloc(128 * GiB);
        CHECK(big && g_used == 160 * GiB, "128 GiB fits beside the cache");
        pool.release(big);
        CHECK(pool.cached_bytes == 128 * GiB && g_used == 128 * GiB, "the 16 GiB blocks were evicted for the 128 GiB one (%zu GiB cached)", pool.cached_bytes >> 30);
        void *again = pool.alloc(128 * GiB);
        CHECK(again == big, "and it is re-used");
        pool.release(again);
        void *s1 = pool.alloc(16 * GiB), *s2 = pool.alloc(16 * GiB);
        pool.release(s1);
        CHECK(pool.cached_bytes == 16 * GiB && g_used == 32 * GiB, "a 16 GiB release evicts the 128 GiB block (%zu GiB cached, %zu used)", pool.cached_bytes >> 30, g_used >> 30);
        pool.release(s2);
        /* allocation failure trims the cache and retries */
        void *huge = pool.alloc(128 * GiB);
        void *more = pool.alloc(32 * GiB);
        CHECK(huge && more && g_used == 160 * GiB, "trim-and-retry");
        bool threw = false;
        try {
            pool.alloc(64 * GiB);
        } catch (const std::runtime_error &) {
            threw = true;
        }
        CHECK(threw, "a real out-of-memory is reported");
        pool.release(huge);
        pool.release(more);
        pool.clear();
        CHECK(g_blocks.empty() && g_used == 0, "clear() frees everything (2)");
    }
    { /* exported blocks with kept peer mappings */
        MemPool pool;
        pool.pin_exported = true;
        void *m = pool.alloc(16 * GiB), *a = pool.alloc(16 * GiB);
        pool.exported.insert(m);
        pool.exported.insert(a);
        g_pinned_by_peer.insert(m); /* the peers keep their mappings after the state vector is gone */
        g_pinned_by_peer.insert(a);
        pool.release(m);
        pool.release(a);
        void *other = pool.alloc(8 * GiB);
        pool.release(other); /* another size class: must not evict the exported ones */
        void *big = pool.alloc(100 * GiB);
        pool.release(big);   /* 100 GiB -> class 128 GiB: over the cap with 32 + 8 cached; only the 8 GiB block may go */
        CHECK(g_blocks.count(m) && g_blocks.count(a), "exported blocks survive eviction");
        bool threw = false;
        try {
            void *p1 = pool.alloc(128 * GiB); /* cached 128 GiB class block */
            void *p2 = pool.alloc(32 * GiB);  /* 170 - 128 - 32 (exported) = 10 GiB free: must fail, NOT free m / a */
            (void)p1, (void)p2;
        } catch (const std::runtime_error &) {
            threw = true;
        }
        CHECK(threw && g_blocks.count(m) && g_blocks.count(a), "trim() on failure leaves exported blocks alone");
        /* the next state vector of the same size gets them back */
        void *m2 = pool.alloc(16 * GiB), *a2 = pool.alloc(16 * GiB);
        CHECK((m2 == m || m2 == a) && (a2 == m || a2 == a) && m2 != a2, "exported blocks are re-used");
        pool.release(m2);
        pool.release(a2);
        /* a sharded state of another size: the ranks drop their mappings, then trim_exported() */
        g_pinned_by_peer.clear();
        const size_t before = g_used;
        pool.trim_exported();
        CHECK(!g_blocks.count(m) && !g_blocks.count(a) && g_used == before - 32 * GiB && pool.exported.empty(), "trim_exported() frees them");
        pool.clear();
        CHECK(g_blocks.empty() && g_used == 0, "clear() frees everything (3)");
    }
    { /* without pinning (ranks sharing a GPU unmap at once) an exported block is an ordinary block */
        MemPool pool;
        void *x[3] = {pool.alloc(16 * GiB), pool.alloc(16 * GiB), pool.alloc(16 * GiB)};
        for (void *p : x) pool.exported.insert(p);
        for (void *p : x) pool.release(p);
        CHECK(pool.cached_bytes == 32 * GiB && pool.exported.size() == 2, "the third block is freed and forgotten (%zu exported)", pool.exported.size());
        pool.clear();
    }
    if (g_failures) {
        std::printf("%d FAILURES\n", g_failures);
        return 1;
    }
    std::printf("pool: %ld mallocs, %ld frees\nALL OK\n", g_mallocs, g_frees);
    return 0;
}
