// TEST INFRASTRUCTURE ONLY -- fiber scheduler behind tests/emul/cuda_emul.h.
#include "cuda_emul.h"

#include <ucontext.h>

thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;

namespace emul {

static thread_local BlockState* g_block = nullptr;
static const size_t kStackBytes = 256 * 1024;

BlockState& block() { return *g_block; }
unsigned char* dyn_smem() { return g_block->dyn_smem.data(); }

#if defined(__x86_64__)
// Minimal cooperative context switch: callee-saved registers + stack pointer.
extern "C" void mgb_emul_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl mgb_emul_switch
.type mgb_emul_switch,@function
mgb_emul_switch:
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
.size mgb_emul_switch,.-mgb_emul_switch
)");

static void fiber_main() {
    BlockState& b = *g_block;
    b.body();
    Fiber& f = b.fibers[b.current];
    f.done = true;
    b.alive--;
    // an exited thread no longer takes part in barriers: release one that is now complete
    if (b.alive > 0 && b.arrived >= b.alive && b.arrived > 0) {
        b.arrived = 0;
        b.gen++;
    }
    void* dummy;
    mgb_emul_switch(&dummy, b.sched_sp);
    std::abort();  // never resumed
}

static void prepare(Fiber& f) {
    if (!f.stack) f.stack = (char*)std::malloc(kStackBytes);
    uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
    void** sp = (void**)(top - 8);
    *sp = nullptr;                   // fake return address of fiber_main
    *(--sp) = (void*)&fiber_main;    // popped by `ret`
    for (int i = 0; i < 6; ++i) *(--sp) = nullptr;
    f.sp = (void*)sp;
    f.done = false;
}

void yield() {
    BlockState& b = *g_block;
    Fiber& f = b.fibers[b.current];
    mgb_emul_switch(&f.sp, b.sched_sp);
}

static void resume(BlockState& b, unsigned i) {
    b.current = i;
    threadIdx.x = b.fibers[i].tid % blockDim.x;
    threadIdx.y = (b.fibers[i].tid / blockDim.x) % blockDim.y;
    threadIdx.z = b.fibers[i].tid / (blockDim.x * blockDim.y);
    mgb_emul_switch(&b.sched_sp, b.fibers[i].sp);
}
#else
#error "the emulator's context switch is written for x86-64"
#endif

void launch(dim3 grid, dim3 bdim, size_t smem_bytes, const std::function<void()>& body) {
    static thread_local BlockState state;
    BlockState* outer = g_block;
    BlockState& b = state;
    g_block = &b;
    gridDim = grid;
    blockDim = bdim;
    unsigned nthreads = bdim.x * bdim.y * bdim.z;
    if (b.fibers.size() < nthreads) b.fibers.resize(nthreads);
    b.warps.assign((nthreads + 31) / 32, WarpState());
    b.dyn_smem.assign(smem_bytes + 64, 0);
    b.body = body;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                b.nthreads = b.alive = nthreads;
                b.arrived = 0;
                for (auto& w : b.warps) { w.arrived = 0; }
                for (unsigned i = 0; i < nthreads; ++i) { b.fibers[i].tid = i; prepare(b.fibers[i]); }
                while (b.alive > 0) {
                    bool progressed = false;
                    for (unsigned i = 0; i < nthreads; ++i) {
                        if (b.fibers[i].done) continue;
                        resume(b, i);
                        progressed = true;
                    }
                    if (!progressed) break;
                }
            }
    g_block = outer;
}

}  // namespace emul
