// glsl_shim.hpp -- just enough of GLSL 4.60 compute (+ KHR_shader_subgroup basic / arithmetic / ballot) in C++17 to
// compile the reference's three shaders from their own text and run them on the CPU.
//
// TEST INFRASTRUCTURE (oracle/_ref): it pins the hand-written restatement oracle/vkrs_oracle.c -- and through it the
// CUDA kernels -- to the reference's own source.  Nothing under vkradixsort_b200/ links, loads or runs this.
//
// How a shader is executed: the shader text (#version / #extension lines dropped, `shared T[N] name;` rewritten to
// GLSL_SHARED(T, N, name); by oracle/Makefile -- every statement of main() is the reference's) becomes the body of a
// C++ struct: interface blocks turn into anonymous structs of pointers, `shared` variables into members (one object
// = one work group), main() into a member function.  Each of the 256 invocations of a work group is a FIBER with its
// own stack running main() on that object; barrier() and the subgroup operations are rendezvous points at which a
// fiber yields to the next one (round robin, lane order), so the execution is deterministic.  Subgroups are 32
// consecutive invocations (SUBGROUP_SIZE 32, multi_radixsort.comp:13).  The shaders only call subgroup operations in
// control flow that is uniform over the subgroup, which the shim asserts.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace glsl {

typedef unsigned int uint;
struct uvec3 {
    uint x, y, z;
};

constexpr int kSubgroup = 32;

// ---- fibers: a minimal x86-64 System V context switch (callee-saved registers + stack pointer) ----
struct Fiber {
    void *sp = nullptr;
    std::vector<unsigned char> stack;
    bool done = true;
};

extern "C" void glsl_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl glsl_switch
.type glsl_switch,@function
glsl_switch:
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
.size glsl_switch,.-glsl_switch
)");

struct Invocation {
    uvec3 global_id{}, local_id{}, workgroup_id{};
    uint subgroup_id = 0, subgroup_invocation_id = 0;
    Fiber fiber;
};

struct Subgroup {
    uint value[kSubgroup];
    uint arrived = 0;     // lanes that have deposited a value for the current operation
    uint generation = 0;  // completed operations
    uint result_sum = 0;
    uint excl[kSubgroup];
};

struct WorkGroupRun {
    std::vector<Invocation> inv;
    std::vector<Subgroup> sub;
    uint barrier_arrived = 0, barrier_generation = 0, live = 0;
    void *scheduler_sp = nullptr;
    Invocation *cur = nullptr;
    void (*entry)(void *) = nullptr;
    void *entry_arg = nullptr;
};

inline WorkGroupRun *&run() {
    static thread_local WorkGroupRun *r = nullptr;
    return r;
}
inline Invocation *cur() { return run()->cur; }

inline void yield() { // back to the scheduler; it resumes the next fiber
    WorkGroupRun *r = run();
    glsl_switch(&r->cur->fiber.sp, r->scheduler_sp);
}

inline void fiber_trampoline() {
    WorkGroupRun *r = run();
    r->entry(r->entry_arg);
    r->cur->fiber.done = true;
    r->live--;
    yield();
    abort(); // a finished fiber is never resumed
}

// Runs `entry(arg)` once per invocation of one work group of `size` invocations (a multiple of 32).
inline void run_workgroup(WorkGroupRun &r, uint size, uint workgroup_id, void (*entry)(void *), void *arg) {
    constexpr size_t kStack = 64 * 1024;
    if (r.inv.size() != size) {
        r.inv.assign(size, Invocation());
        for (auto &i : r.inv) i.fiber.stack.resize(kStack);
        r.sub.assign(size / kSubgroup, Subgroup());
    }
    for (auto &s : r.sub) s = Subgroup();
    r.barrier_arrived = r.barrier_generation = 0;
    r.live = size;
    r.entry = entry;
    r.entry_arg = arg;
    for (uint l = 0; l < size; ++l) {
        Invocation &i = r.inv[l];
        i.local_id = {l, 0, 0};
        i.workgroup_id = {workgroup_id, 0, 0};
        i.global_id = {workgroup_id * size + l, 0, 0};
        i.subgroup_id = l / kSubgroup;
        i.subgroup_invocation_id = l % kSubgroup;
        i.fiber.done = false;
        // initial frame: six callee-saved registers (zero) + the return address = the trampoline; the stack pointer
        // is 16-byte aligned + 8 when the trampoline starts, as after a call
        auto top = reinterpret_cast<uintptr_t>(i.fiber.stack.data() + kStack) & ~uintptr_t(15);
        void **sp = reinterpret_cast<void **>(top) - 1; // keeps (rsp + 8) % 16 == 0 at entry
        *--sp = reinterpret_cast<void *>(&fiber_trampoline);
        for (int k = 0; k < 6; ++k) *--sp = nullptr;
        i.fiber.sp = sp;
    }
    WorkGroupRun *prev = run();
    run() = &r;
    while (r.live > 0) {
        for (uint l = 0; l < size; ++l) {
            if (r.inv[l].fiber.done) continue;
            r.cur = &r.inv[l];
            glsl_switch(&r.scheduler_sp, r.inv[l].fiber.sp);
        }
    }
    run() = prev;
}

// ---- barrier(): every invocation of the work group that has not returned yet ----
inline void barrier() {
    WorkGroupRun *r = run();
    const uint gen = r->barrier_generation;
    if (++r->barrier_arrived == r->live) {
        r->barrier_arrived = 0;
        r->barrier_generation++;
        return;
    }
    while (r->barrier_generation == gen) {
        yield();
        // an invocation that returns from main() while others wait would deadlock a real GPU as well; the shaders never do that
    }
}

// ---- subgroup operations (all 32 lanes of the subgroup take part: uniform control flow) ----
inline Subgroup &subgroup_collect(uint v) {
    WorkGroupRun *r = run();
    Invocation *me = r->cur;
    Subgroup &s = r->sub[me->subgroup_id];
    const uint gen = s.generation;
    s.value[me->subgroup_invocation_id] = v;
    if (++s.arrived == kSubgroup) {
        uint acc = 0;
        for (int l = 0; l < kSubgroup; ++l) {
            s.excl[l] = acc;
            acc += s.value[l];
        }
        s.result_sum = acc;
        s.arrived = 0;
        s.generation++;
    } else {
        while (s.generation == gen) yield();
    }
    return s;
}
// The results of one operation are read before any lane can complete the next one: a lane only deposits into the
// next operation after it has returned from this one, and completion needs all 32 deposits.
inline uint subgroupAdd(uint v) { return subgroup_collect(v).result_sum; }
inline uint subgroupExclusiveAdd(uint v) {
    const uint lane = cur()->subgroup_invocation_id;
    return subgroup_collect(v).excl[lane];
}
inline uint subgroupBroadcast(uint v, uint id) {
    Subgroup &s = subgroup_collect(v);
    assert(id < (uint) kSubgroup);
    return s.value[id];
}
inline bool subgroupElect() { return cur()->subgroup_invocation_id == 0; } // lowest active lane; all lanes are active here

inline uint atomicAdd(uint &mem, uint v) {
    const uint old = mem;
    mem = old + v;
    return old;
}
inline uint bitCount(uint v) { return (uint) __builtin_popcount(v); }

} // namespace glsl

// ---- the names a shader's text uses ----
using glsl::uint;
#define layout(...)
#define in
#define uniform struct
#define buffer struct
#define GLSL_SHARED(T, N, name) T name[(N) + 64] = {} /* + slack: multi_radixsort.comp:74 reads sums[8..31] */
#define gl_GlobalInvocationID (glsl::cur()->global_id)
#define gl_LocalInvocationID (glsl::cur()->local_id)
#define gl_WorkGroupID (glsl::cur()->workgroup_id)
#define gl_SubgroupID (glsl::cur()->subgroup_id)
#define gl_SubgroupInvocationID (glsl::cur()->subgroup_invocation_id)
#define barrier glsl::barrier
#define subgroupAdd glsl::subgroupAdd
#define subgroupExclusiveAdd glsl::subgroupExclusiveAdd
#define subgroupBroadcast glsl::subgroupBroadcast
#define subgroupElect glsl::subgroupElect
#define atomicAdd glsl::atomicAdd
#define bitCount glsl::bitCount
// interface block names vanish (the blocks become anonymous structs whose members are the shader's globals) ...
#define PushConstants
#define elements_in
#define elements_out
#define histograms
// ... and an unsized array member `uint g_x[];` becomes a pointer to an array of unknown bound
#define g_elements_in (*g_elements_in_p)
#define g_elements_out (*g_elements_out_p)
#define g_histograms (*g_histograms_p)
