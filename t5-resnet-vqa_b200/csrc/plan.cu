// Launch plans: record fully-resolved kernel launches once, replay them with one host call
// (optionally as an instantiated CUDA graph).  See include/vqa_b200.h "launch plans".
#include "../../include/vqa_b200.h"
#include "common.cuh"

#include <string>
#include <vector>

using namespace vqa;

extern "C" {

void* vqa_plan_create(void) { return new Plan(); }

int vqa_plan_destroy(void* plan) {
  if (plan == nullptr) return 0;
  Plan* p = static_cast<Plan*>(plan);
  if (p->exec) cudaGraphExecDestroy(p->exec);
  if (p->graph) cudaGraphDestroy(p->graph);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  for (cudaEvent_t e : p->marks) if (e) cudaEventDestroy(e);
  if (p->side) cudaStreamDestroy(p->side);
  delete p;
  return 0;
}

int vqa_plan_size(void* plan) {
  return plan ? static_cast<int>(static_cast<Plan*>(plan)->ops.size()) : 0;
}

int vqa_plan_set_lane(void* plan, int lane) {
  if (plan == nullptr || (lane != 0 && lane != 1)) { set_last_error("plan_set_lane: bad arguments"); return -1; }
  static_cast<Plan*>(plan)->cur_lane = lane;
  return 0;
}

int vqa_plan_fork(void* plan) {
  if (plan == nullptr) return -1;
  Plan* p = static_cast<Plan*>(plan);
  p->steps.push_back({PLAN_FORK, -1});
  return 0;
}

int vqa_plan_join(void* plan) {
  if (plan == nullptr) return -1;
  Plan* p = static_cast<Plan*>(plan);
  p->steps.push_back({PLAN_JOIN, -1});
  p->lane1_open = false;
  return 0;
}

int vqa_plan_mark(void* plan) {
  if (plan == nullptr) return -1;
  Plan* p = static_cast<Plan*>(plan);
  p->marks.push_back(nullptr);
  const int id = static_cast<int>(p->marks.size()) - 1;
  p->steps.push_back({PLAN_MARK, id});
  return id;
}

int vqa_plan_wait(void* plan, int mark) {
  if (plan == nullptr) return -1;
  Plan* p = static_cast<Plan*>(plan);
  if (mark < 0 || mark >= static_cast<int>(p->marks.size())) { set_last_error("plan_wait: unknown mark %d", mark); return -1; }
  p->steps.push_back({PLAN_WAIT, mark});
  return 0;
}

static int ensure_side(Plan* p) {
  if (p->side != nullptr) return 0;
  cudaError_t e = cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming);
  if (e != cudaSuccess) { set_last_error("plan side stream: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

static int replay(Plan* p, cudaStream_t s) {
  bool forked = false;   // lane 1 has been ordered after lane 0 at least once in this replay
  bool open = false;     // lane 1 holds work lane 0 has not waited for
  std::vector<char> marked(p->marks.size(), 0);   // marks recorded during THIS replay
  for (const Plan::Step& st : p->steps) {
    if (st.kind == PLAN_LAUNCH) {
      cudaStream_t target = s;
      if (p->lanes[st.op] == 1) {
        int r = ensure_side(p);
        if (r) return r;
        if (!forked) {   // implicit fork: lane 1 never runs ahead of the point where the plan started
          cudaEventRecord(p->ev_fork, s);
          cudaStreamWaitEvent(p->side, p->ev_fork, 0);
          forked = true;
        }
        target = p->side;
        open = true;
      }
      int r = p->ops[st.op](target);
      if (r) return r;
    } else if (st.kind == PLAN_FORK) {
      int r = ensure_side(p);
      if (r) return r;
      cudaEventRecord(p->ev_fork, s);
      cudaStreamWaitEvent(p->side, p->ev_fork, 0);
      forked = true;
    } else if (st.kind == PLAN_MARK) {
      if (!forked) continue;   // lane 1 has done nothing yet: nothing to wait for
      if (p->marks[st.op] == nullptr &&
          cudaEventCreateWithFlags(&p->marks[st.op], cudaEventDisableTiming) != cudaSuccess) {
        set_last_error("plan mark: cannot create event");
        return -1;
      }
      cudaEventRecord(p->marks[st.op], p->side);
      marked[st.op] = 1;
    } else if (st.kind == PLAN_WAIT) {
      if (marked[st.op]) cudaStreamWaitEvent(s, p->marks[st.op], 0);
    } else if (open) {  // PLAN_JOIN
      cudaEventRecord(p->ev_join, p->side);
      cudaStreamWaitEvent(s, p->ev_join, 0);
      open = false;
    }
  }
  if (open) {  // a plan always ends joined: its caller only knows lane 0's stream
    cudaEventRecord(p->ev_join, p->side);
    cudaStreamWaitEvent(s, p->ev_join, 0);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error("plan replay: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

int vqa_plan_run(void* plan, void* stream) {
  if (plan == nullptr) { set_last_error("plan_run: null plan"); return -1; }
  Plan* p = static_cast<Plan*>(plan);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->exec) {
    cudaError_t e = cudaGraphLaunch(p->exec, s);
    if (e != cudaSuccess) {
      set_last_error("cudaGraphLaunch: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    return 0;
  }
  return replay(p, s);
}

int vqa_plan_capture_graph(void* plan, void* stream) {
  if (plan == nullptr) { set_last_error("plan_capture: null plan"); return -1; }
  Plan* p = static_cast<Plan*>(plan);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->exec) { cudaGraphExecDestroy(p->exec); p->exec = nullptr; }
  if (p->graph) { cudaGraphDestroy(p->graph); p->graph = nullptr; }
  cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    set_last_error("cudaStreamBeginCapture: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  int r = replay(p, s);
  cudaGraph_t g = nullptr;
  e = cudaStreamEndCapture(s, &g);
  if (r) { if (g) cudaGraphDestroy(g); return r; }
  if (e != cudaSuccess) {
    set_last_error("cudaStreamEndCapture: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  e = cudaGraphInstantiate(&p->exec, g, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(g);
    p->exec = nullptr;
    set_last_error("cudaGraphInstantiate: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  p->graph = g;
  return 0;
}

// ---- per-launch timing of a recorded plan (bench.py's roofline numbers) --------------------------------
// Replays the plan eagerly with one CUDA event between consecutive launches; a short device-side spin in
// front lets the host queue everything before the first launch starts, so the event-to-event times are the
// launches' back-to-back device durations, not host enqueue gaps.  ms_out has vqa_plan_size(plan) entries.
__global__ void vqa_spin_kernel(long long ns) {
  pdl_grid_sync();
  long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < ns);
}

int vqa_plan_profile(void* plan, void* stream, float* ms_out, int spin_us) {
  if (plan == nullptr) { set_last_error("plan_profile: null plan"); return -1; }
  Plan* p = static_cast<Plan*>(plan);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = p->ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) cudaEventCreate(&e);
  int rc = 0;
  if (spin_us > 0) launch_pdl(vqa_spin_kernel, dim3(1), dim3(1), 0, s, static_cast<long long>(spin_us) * 1000);
  for (size_t i = 0; i < n && rc == 0; ++i) {
    cudaEventRecord(ev[i], s);
    rc = p->ops[i](s);
  }
  cudaEventRecord(ev[n], s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc == 0 && e != cudaSuccess) {
    set_last_error("plan_profile: %s", cudaGetErrorString(e));
    rc = static_cast<int>(e);
  }
  if (rc == 0)
    for (size_t i = 0; i < n; ++i) cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
  for (auto& ev_i : ev) cudaEventDestroy(ev_i);
  return rc;
}

// Back-to-back device time of a family of launches: replays only the ops whose name is listed in `names` (comma
// separated, e.g. "gemm,conv,conv_wgrad") `reps` times on `stream` between ONE pair of events, programmatic dependent
// launch on, nothing in between - the kernels' durations as they run inside the step, without the event-to-event gap
// the per-launch table above pays for every launch.  Data dependencies are ignored (timing only; run it last).
int vqa_plan_time_ops(void* plan, void* stream, const char* names, int reps, float* ms_per_rep, double* flops_per_rep,
                      int* launches_per_rep) {
  if (plan == nullptr || names == nullptr || reps < 1) { set_last_error("plan_time_ops: bad arguments"); return -1; }
  Plan* p = static_cast<Plan*>(plan);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const std::string list = std::string(",") + names + ",";
  std::vector<size_t> sel;
  double fl = 0.0;
  for (size_t i = 0; i < p->ops.size(); ++i) {
    const std::string key = std::string(",") + p->notes[i].name + ",";
    if (list.find(key) != std::string::npos) { sel.push_back(i); fl += p->notes[i].flops; }
  }
  *flops_per_rep = fl;
  *launches_per_rep = static_cast<int>(sel.size());
  *ms_per_rep = 0.f;
  if (sel.empty()) return 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = 0;
  launch_pdl(vqa_spin_kernel, dim3(1), dim3(1), 0, s, 2000000LL);   // let the host queue ahead of the device
  for (size_t i : sel) { rc = p->ops[i](s); if (rc) break; }          // one untimed pass (caches, attributes)
  cudaEventRecord(e0, s);
  for (int r = 0; r < reps && rc == 0; ++r)
    for (size_t i : sel) { rc = p->ops[i](s); if (rc) break; }
  cudaEventRecord(e1, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc == 0 && e != cudaSuccess) { set_last_error("plan_time_ops: %s", cudaGetErrorString(e)); rc = static_cast<int>(e); }
  if (rc == 0) { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *ms_per_rep = ms / reps; }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

int vqa_plan_op_info(void* plan, int i, const char** name, double* flops, double* bytes) {
  if (plan == nullptr) return -1;
  Plan* p = static_cast<Plan*>(plan);
  if (i < 0 || i >= static_cast<int>(p->notes.size())) { set_last_error("plan_op_info: index out of range"); return -1; }
  *name = p->notes[i].name; *flops = p->notes[i].flops; *bytes = p->notes[i].bytes;
  return 0;
}

}  // extern "C"
