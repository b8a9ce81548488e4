// TEST INFRASTRUCTURE: run-time of tests/emul/cuda_shim.h (user-level lane contexts, x86-64 System V).
#include "cuda_shim.h"

#undef __noinline__

thread_local shim::WarpRt* shim::rt = nullptr;
thread_local int shim::lane = 0;
thread_local shim::Dim threadIdx, blockIdx, blockDim;

// void shim_switch(void** save_sp, void* load_sp): saves the callee-saved registers of the running context on
// its stack, stores the stack pointer, loads the other context's stack pointer and registers, returns into it
asm(R"(
.text
.globl shim_switch
.type shim_switch,@function
shim_switch:
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
.size shim_switch,.-shim_switch
)");

namespace shim {

static void lane_main() {
  WarpRt* w = rt;
  const int me = w->cur;
  w->body(static_cast<int>(w->base_tid) + me);
  w = rt;
  w->done[me] = true;
  w->n_done += 1;
  int next = (me + 1) & 31;
  for (int k = 0; k < 32 && w->done[next]; ++k) next = (next + 1) & 31;
  to_lane(w->n_done == 32 ? -1 : next);
  std::abort();   // a finished lane is never resumed
}

void to_lane(int next) {
  WarpRt* w = rt;
  const int me = w->cur;
  if (next < 0) {   // all lanes are done: back to the warp's thread
    void* dummy;
    shim_switch(&dummy, w->main_sp);
    return;
  }
  if (next == me) return;
  w->cur = next;
  lane = next;
  threadIdx.x = w->base_tid + static_cast<unsigned>(next);
  shim_switch(&w->sp[me], w->sp[next]);
  // resumed: whoever switched to me has set cur / lane / threadIdx for me
}

void run_warp(WarpRt* w) {
  rt = w;
  w->stacks = static_cast<char*>(std::aligned_alloc(4096, 32 * kStackBytes));
  for (int l = 0; l < 32; ++l) {
    w->done[l] = false;
    w->ncoll[l] = 0;
    // initial frame: six callee-saved registers (zero), then the return address = lane_main at a 16-byte aligned slot
    uintptr_t top = reinterpret_cast<uintptr_t>(w->stacks + (l + 1) * kStackBytes);
    top &= ~static_cast<uintptr_t>(15);
    void** sp = reinterpret_cast<void**>(top);
    *--sp = nullptr;                                   // after `ret`: rsp % 16 == 8, as at any function entry
    *--sp = reinterpret_cast<void*>(&lane_main);       // return address, 16-byte aligned slot
    for (int r = 0; r < 6; ++r) *--sp = nullptr;
    w->sp[l] = sp;
  }
  w->cur = 0;
  w->n_done = 0;
  lane = 0;
  threadIdx.x = w->base_tid;
  shim_switch(&w->main_sp, w->sp[0]);
  std::free(w->stacks);
  w->stacks = nullptr;
  rt = nullptr;
}

}  // namespace shim
