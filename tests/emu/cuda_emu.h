// Minimal host emulation of the CUDA execution model for kernels that use only threadIdx / blockIdx, __shared__ arrays and
// __syncthreads(): every thread of a block is a pthread, __syncthreads() is a pthread barrier, blocks run one after the other
// (a __shared__ array is a function-local static, shared by the block's threads).  Compiled with -fsanitize=address or
// -fsanitize=thread this checks every global / shared index the kernel forms and every missing barrier at the sizes the test drives.
// Test infrastructure only.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <functional>
#include <vector>

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

static thread_local dim3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static pthread_barrier_t emu_barrier;

// warp-level primitives: the 32 lanes of a warp are 32 pthreads; a convergent warp operation is two waits on the warp's barrier
// around an exchange buffer (the kernels this header serves call them from warp-uniform control flow only)
constexpr int EMU_MAX_WARPS = 32;
static pthread_barrier_t emu_warp_barrier[EMU_MAX_WARPS];
static uint32_t emu_xchg[EMU_MAX_WARPS][32];
static float *emu_dyn_smem_ptr = nullptr;      // "extern __shared__": one buffer per launch, shared by the block's threads
static inline float *emu_dyn_smem_float() { return emu_dyn_smem_ptr; }
static inline int emu_linear_tid() { return (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)); }

#define __global__
#define __shared__ static
#define __launch_bounds__(...)
#define __device__
#define __forceinline__ inline

#define __host__
static inline void __syncthreads() { pthread_barrier_wait(&emu_barrier); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu_warp_barrier[emu_linear_tid() >> 5]); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    const int w = emu_linear_tid() >> 5, l = emu_linear_tid() & 31;
    memcpy(&emu_xchg[w][l], &v, 4);
    pthread_barrier_wait(&emu_warp_barrier[w]);
    T r;
    memcpy(&r, &emu_xchg[w][l ^ lane_mask], 4);
    pthread_barrier_wait(&emu_warp_barrier[w]);
    return r;
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

struct EmuThread {
    dim3 t, b;
    const std::function<void()> *body;
};

static void *emu_thread_main(void *p) {
    EmuThread *e = static_cast<EmuThread *>(p);
    threadIdx = e->t;
    blockIdx = e->b;
    (*e->body)();
    return nullptr;
}

// launch<<<grid, block>>>: body() is the kernel call with its arguments bound
static inline void emu_launch(dim3 grid, dim3 block, const std::function<void()> &body) {
    gridDim = grid;
    blockDim = block;
    const unsigned nt = block.x * block.y * block.z;
    pthread_barrier_init(&emu_barrier, nullptr, nt);
    const unsigned n_warps = (nt + 31) / 32;
    for (unsigned w = 0; w < n_warps && w < (unsigned)EMU_MAX_WARPS; ++w)
        pthread_barrier_init(&emu_warp_barrier[w], nullptr, (w + 1) * 32 <= nt ? 32 : nt - w * 32);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                std::vector<pthread_t> th(nt);
                std::vector<EmuThread> arg(nt);
                unsigned n = 0;
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx, ++n) {
                            arg[n] = EmuThread{dim3(tx, ty, tz), dim3(bx, by, bz), &body};
                            pthread_create(&th[n], nullptr, emu_thread_main, &arg[n]);
                        }
                for (unsigned i = 0; i < nt; ++i) pthread_join(th[i], nullptr);
            }
    pthread_barrier_destroy(&emu_barrier);
    for (unsigned w = 0; w < n_warps && w < (unsigned)EMU_MAX_WARPS; ++w) pthread_barrier_destroy(&emu_warp_barrier[w]);
}
