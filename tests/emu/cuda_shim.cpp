// cuda_shim.cpp -- fiber scheduler behind tests/emu/cuda_shim.h (TEST INFRASTRUCTURE).
#include "cuda_shim.h"

#include <stdio.h>

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
std::vector<unsigned char> g_stacks;
std::vector<unsigned char> g_smem;
const std::function<void()> *g_body = nullptr;
unsigned g_cur = 0;

void trampoline()
{
    (*g_body)();
    g_fibers[g_cur].done = true;
    swapcontext(&g_fibers[g_cur].ctx, &g_main);
}
}   // namespace

unsigned char *block_smem() { return g_smem.data(); }

void sync() { swapcontext(&g_fibers[g_cur].ctx, &g_main); }

void run_grid(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body)
{
    g_body = &body;
    gridDim.x = grid;
    blockDim.x = block;
    if (g_stacks.size() < (size_t)block * STACK) g_stacks.resize((size_t)block * STACK);
    g_smem.assign(smem + 16, 0xA5);   // poison: kernels must not read uninitialised shared memory
    g_fibers.resize(block);
    for (unsigned b = 0; b < grid; b++) {
        blockIdx.x = b;
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = g_fibers[t];
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t)t * STACK;
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
        }
        for (;;) {
            unsigned alive = 0, finished = 0;
            for (unsigned t = 0; t < block; t++) {
                if (g_fibers[t].done) continue;
                alive++;
                g_cur = t;
                threadIdx.x = t;
                swapcontext(&g_main, &g_fibers[t].ctx);
                if (g_fibers[t].done) finished++;
            }
            if (alive == 0) break;
            if (finished != 0 && finished != alive) {
                fprintf(stderr, "emu: divergent barrier in block %u (%u of %u threads exited)\n", b, finished, alive);
                abort();
            }
        }
    }
}
}   // namespace emu
