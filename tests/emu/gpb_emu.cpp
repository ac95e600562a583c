// tests/emu/gpb_emu.cpp -- TEST INFRASTRUCTURE.  Fiber-based CPU emulator for the
// geepee_b200 kernels (built with -DGPB_CPU_EMU, see geepee_b200/csrc/gpb_rt.cuh).
// Every CUDA thread of a block runs as a ucontext coroutine; __syncthreads and warp
// shuffles yield to a scheduler that releases them when all participants arrived.
// Blocks run one after another.  GPB_EMU_REVERSE=1 runs threads in descending order
// (a cheap way to expose missing-barrier hazards).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <stdint.h>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace gpb_emu {

dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
unsigned char* dyn_smem = nullptr;

enum { READY = 0, AT_BARRIER = 1, AT_SHFL = 2, DONE = 3, AT_GROUP = 4 };
struct Fiber {
    ucontext_t ctx;
    unsigned char* stack;
    int state, mask;
    double val, val2, res, res2;
};
static const size_t kStack = 256 * 1024;
static std::vector<Fiber> fibers;
static std::vector<unsigned char> smem_buf;
static ucontext_t sched;
static int cur = -1, nthreads = 0;
static long blk = -1, nblk = 0;
static void (*g_tramp)(void*) = nullptr;
static void* g_ctx = nullptr;

void launch_begin(dim3 grid, dim3 block, size_t smem) {
    t_gridDim = grid;
    t_blockDim = block;
    nthreads = (int)(block.x * block.y * block.z);
    nblk = (long)grid.x * grid.y * grid.z;
    blk = -1;
    smem_buf.assign(smem + 64, 0);
    dyn_smem = (unsigned char*)(((uintptr_t)smem_buf.data() + 63) & ~(uintptr_t)63);
    while ((int)fibers.size() < nthreads) {
        Fiber f;
        memset(&f, 0, sizeof(f));
        f.stack = (unsigned char*)malloc(kStack);
        fibers.push_back(f);
    }
}

bool launch_next_block() {
    blk++;
    if (blk >= nblk) return false;
    t_blockIdx.x = (unsigned)(blk % t_gridDim.x);
    t_blockIdx.y = (unsigned)((blk / t_gridDim.x) % t_gridDim.y);
    t_blockIdx.z = (unsigned)(blk / ((long)t_gridDim.x * t_gridDim.y));
    return true;
}

static void fiber_entry() {
    g_tramp(g_ctx);
    fibers[cur].state = DONE;
    swapcontext(&fibers[cur].ctx, &sched);
}

void barrier() {
    fibers[cur].state = AT_BARRIER;
    swapcontext(&fibers[cur].ctx, &sched);
}

void group_barrier(int n) {
    fibers[cur].state = AT_GROUP;
    fibers[cur].mask = n;
    swapcontext(&fibers[cur].ctx, &sched);
}

double shfl_xor_f64(double v, int m) {
    Fiber& f = fibers[cur];
    f.state = AT_SHFL;
    f.val = v;
    f.mask = m;
    swapcontext(&f.ctx, &sched);
    return fibers[cur].res;
}

double shfl_idx_f64(double v, int src) {
    Fiber& f = fibers[cur];
    f.state = AT_SHFL;
    f.val = v;
    f.mask = 0x100 | (src & 31);   // bit 8: indexed read instead of xor
    swapcontext(&f.ctx, &sched);
    return fibers[cur].res;
}

// DMMA.8x8x4 in one warp rendezvous: every lane deposits its A and B fragment element, the
// scheduler computes the lane's two C increments (fragment layout of the PTX instruction).
void dmma_f64(double a, double b, double* d0, double* d1) {
    Fiber& f = fibers[cur];
    f.state = AT_SHFL;
    f.val = a;
    f.val2 = b;
    f.mask = 0x200;
    swapcontext(&f.ctx, &sched);
    *d0 = fibers[cur].res;
    *d1 = fibers[cur].res2;
}

void run_block(void (*tramp)(void*), void* ctx) {
    g_tramp = tramp;
    g_ctx = ctx;
    static int reverse = -1;
    if (reverse < 0) reverse = getenv("GPB_EMU_REVERSE") ? 1 : 0;
    for (int i = 0; i < nthreads; i++) {
        Fiber& f = fibers[i];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        f.state = READY;
    }
    for (;;) {
        bool progressed = false;
        for (int k = 0; k < nthreads; k++) {
            int i = reverse ? nthreads - 1 - k : k;
            if (fibers[i].state != READY) continue;
            cur = i;
            t_threadIdx = dim3((unsigned)i, 0, 0);
            swapcontext(&sched, &fibers[i].ctx);
            progressed = true;
        }
        // warp exchanges
        for (int w0 = 0; w0 < nthreads; w0 += 32) {
            int w1 = w0 + 32 < nthreads ? w0 + 32 : nthreads;
            bool all = true, any = false;
            for (int i = w0; i < w1; i++) {
                if (fibers[i].state == AT_SHFL) any = true;
                else if (fibers[i].state != DONE) all = false;
            }
            if (!(all && any)) continue;
            for (int i = w0; i < w1; i++)
                if (fibers[i].state == AT_SHFL && (fibers[i].mask & 0x200)) {
                    const int l = i - w0, g = l >> 2, t = l & 3;
                    double s0 = 0, s1 = 0;
                    for (int k = 0; k < 4; k++) {
                        const double ak = fibers[w0 + g * 4 + k].val;
                        s0 += ak * fibers[w0 + (2 * t) * 4 + k].val2;
                        s1 += ak * fibers[w0 + (2 * t + 1) * 4 + k].val2;
                    }
                    fibers[i].res = s0;
                    fibers[i].res2 = s1;
                } else if (fibers[i].state == AT_SHFL) {
                    int p = (fibers[i].mask & 0x100) ? w0 + (fibers[i].mask & 31)
                                                     : w0 + (((i - w0) ^ fibers[i].mask) & 31);
                    fibers[i].res = (p < w1 && fibers[p].state == AT_SHFL) ? fibers[p].val : fibers[i].val;
                }
            for (int i = w0; i < w1; i++)
                if (fibers[i].state == AT_SHFL) fibers[i].state = READY;
            progressed = true;
        }
        // named barrier over the first n threads
        {
            int n = 0;
            for (int i = 0; i < nthreads; i++)
                if (fibers[i].state == AT_GROUP) { n = fibers[i].mask; break; }
            if (n > 0) {
                bool ok = true;
                for (int i = 0; i < n && i < nthreads; i++)
                    if (fibers[i].state != AT_GROUP) ok = false;
                if (ok) {
                    for (int i = 0; i < n && i < nthreads; i++) fibers[i].state = READY;
                    progressed = true;
                }
            }
        }
        // block barrier
        bool all = true, any = false, alldone = true;
        for (int i = 0; i < nthreads; i++) {
            if (fibers[i].state != DONE) alldone = false;
            if (fibers[i].state == AT_BARRIER) any = true;
            else if (fibers[i].state != DONE) all = false;
        }
        if (alldone) break;
        if (all && any) {
            for (int i = 0; i < nthreads; i++)
                if (fibers[i].state == AT_BARRIER) fibers[i].state = READY;
            progressed = true;
        }
        if (!progressed) {
            fprintf(stderr, "gpb_emu: deadlock (divergent barrier or shuffle) in block %ld\n", blk);
            abort();
        }
    }
}

}  // namespace gpb_emu
