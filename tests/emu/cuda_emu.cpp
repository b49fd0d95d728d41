// tests/emu/cuda_emu.cpp -- fiber scheduler of the CUDA logic emulator.  TEST INFRASTRUCTURE ONLY
// (see cuda_emu.h).
#include "cuda_emu.h"

#include <sys/mman.h>

namespace emu {

Fiber *cur = nullptr;
void *sched_sp = nullptr;
dim3 g_blockIdx, g_blockDim, g_gridDim;
int g_live = 0, g_bar_count = 0, g_bar_gen = 0;
std::vector<WarpSync> g_warps;
unsigned char *g_dyn_smem = nullptr;
const std::function<void()> *g_body = nullptr;
long g_spin_guard = 0;

static std::vector<Fiber> g_fibers;
static std::vector<char *> g_stacks;
static std::vector<unsigned char> g_smem_store;

static void fiber_exit_barrier_fixup()
{
    // an exited thread no longer takes part in block barriers
    if (g_bar_count > 0 && g_bar_count == g_live) {
        g_bar_count = 0;
        g_bar_gen++;
    }
}

extern "C" void emu_trampoline()
{
    (*g_body)();
    cur->done = true;
    g_live--;
    fiber_exit_barrier_fixup();
    emu_switch(&cur->sp, sched_sp);
    abort();  // never resumed
}

static char *get_stack(size_t i)
{
    while (g_stacks.size() <= i) {
        void *p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) {
            perror("emu: mmap stack");
            abort();
        }
        g_stacks.push_back((char *)p);
    }
    return g_stacks[i];
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    size_t nthreads = (size_t)block.x * block.y * block.z;
    assert(block.y == 1 && block.z == 1 && "emulator supports 1-D blocks");
    assert(nthreads % 32 == 0 && nthreads <= 1024);
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) return;
    g_fibers.assign(nthreads, Fiber());
    g_smem_store.assign(smem + 16, 0xCD);
    g_dyn_smem = (unsigned char *)(((uintptr_t)g_smem_store.data() + 15) & ~(uintptr_t)15);
    g_body = &body;
    g_gridDim = grid;
    g_blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = dim3(bx, by, bz);
                g_live = (int)nthreads;
                g_bar_count = 0;
                g_bar_gen = 0;
                g_spin_guard = 0;
                g_warps.assign(nthreads / 32, WarpSync());
                for (size_t t = 0; t < nthreads; ++t) {
                    Fiber &f = g_fibers[t];
                    f.done = false;
                    f.tid = dim3((unsigned)t, 0, 0);
                    f.stack = get_stack(t);
                    uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
                    uint64_t *sp = (uint64_t *)top;
                    *--sp = 0;                              // fake return address (keeps ABI alignment)
                    *--sp = (uint64_t)(uintptr_t)&emu_trampoline;
                    for (int r = 0; r < 6; ++r) *--sp = 0;  // rbp rbx r12 r13 r14 r15
                    f.sp = sp;
                }
                size_t remaining = nthreads;
                while (remaining) {
                    for (size_t t = 0; t < nthreads; ++t) {
                        Fiber &f = g_fibers[t];
                        if (f.done) continue;
                        cur = &f;
                        emu_switch(&sched_sp, f.sp);
                        if (f.done) remaining--;
                    }
                }
            }
    cur = nullptr;
}

}  // namespace emu

asm(R"(
    .text
    .globl emu_switch
    .type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size emu_switch, .-emu_switch
)");
