// cuda_shim.cpp -- fiber scheduler behind tests/emu/cuda_shim.h (TEST INFRASTRUCTURE).
#include "cuda_shim.h"

#include <mutex>

#include <stdio.h>
#include <sys/mman.h>

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
namespace {
constexpr size_t STACK = 256 * 1024;
struct Fiber {
    ucontext_t ctx;
    bool done = false;
};
ucontext_t g_main;
std::vector<Fiber> g_fibers;
unsigned char *g_stacks = nullptr;   // mmap'ed, untouched pages cost nothing
size_t g_stacks_size = 0;
std::vector<unsigned char> g_smem;
size_t g_smem_stride = 0;
const std::function<void()> *g_body = nullptr;
unsigned g_cur = 0, g_block = 1;

void trampoline()
{
    (*g_body)();
    g_fibers[g_cur].done = true;
    swapcontext(&g_fibers[g_cur].ctx, &g_main);
}
}   // namespace

unsigned char *block_smem() { return g_smem.data() + (size_t)(g_cur / g_block) * g_smem_stride; }
unsigned char *cluster_smem(unsigned rank) { return g_smem.data() + (size_t)rank * g_smem_stride; }
unsigned cluster_rank() { return g_cur / g_block; }

// every barrier (block-wide, sub-block or cluster-wide) yields to the scheduler, which resumes a fiber only after
// all fibers of the cluster have yielded: a superset of the synchronisation the kernels ask for
void sync() { swapcontext(&g_fibers[g_cur].ctx, &g_main); }
void cluster_sync() { sync(); }

// the `cluster` consecutive blocks of a thread-block cluster run together (their fibers share one scheduler round)
// one grid at a time: the fiber scheduler and the emulated thread indices are process-wide, while the backend issues work
// for every "GPU" from its own host thread
static std::mutex g_grid_mtx;
void run_grid(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body, unsigned cluster)
{
    std::lock_guard<std::mutex> lock(g_grid_mtx);
    g_body = &body;
    gridDim.x = grid;
    blockDim.x = block;
    g_block = block;
    const unsigned nf = block * cluster;
    if (g_stacks_size < (size_t)nf * STACK) {
        if (g_stacks) munmap(g_stacks, g_stacks_size);
        g_stacks_size = (size_t)nf * STACK;
        g_stacks = (unsigned char *)mmap(nullptr, g_stacks_size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_stacks == MAP_FAILED) { perror("emu: mmap"); abort(); }
    }
    g_smem_stride = (smem + 16 + 127) / 128 * 128;
    g_fibers.resize(nf);
    if (grid % cluster) { fprintf(stderr, "emu: grid %u is not a multiple of the cluster size %u\n", grid, cluster); abort(); }
    for (unsigned b0 = 0; b0 < grid; b0 += cluster) {
        g_smem.assign(g_smem_stride * cluster, 0xA5);   // poison: kernels must not read uninitialised shared memory
        for (unsigned t = 0; t < nf; t++) {
            Fiber &f = g_fibers[t];
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = g_stacks + (size_t)t * STACK;
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
        }
        for (;;) {
            unsigned alive = 0, finished = 0;
            for (unsigned t = 0; t < nf; t++) {
                if (g_fibers[t].done) continue;
                alive++;
                g_cur = t;
                threadIdx.x = t % block;
                blockIdx.x = b0 + t / block;
                swapcontext(&g_main, &g_fibers[t].ctx);
                if (g_fibers[t].done) finished++;
            }
            if (alive == 0) break;
            if (finished != 0 && finished != alive) {
                fprintf(stderr, "emu: divergent barrier in cluster at block %u (%u of %u threads exited)\n", b0, finished, alive);
                abort();
            }
        }
    }
}
}   // namespace emu
