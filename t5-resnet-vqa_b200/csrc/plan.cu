// Launch plans: record fully-resolved kernel launches once, replay them with one host call
// (optionally as an instantiated CUDA graph).  See include/vqa_b200.h "launch plans".
#include "../../include/vqa_b200.h"
#include "common.cuh"

using namespace vqa;

extern "C" {

void* vqa_plan_create(void) { return new Plan(); }

int vqa_plan_destroy(void* plan) {
  if (plan == nullptr) return 0;
  Plan* p = static_cast<Plan*>(plan);
  if (p->exec) cudaGraphExecDestroy(p->exec);
  if (p->graph) cudaGraphDestroy(p->graph);
  delete p;
  return 0;
}

int vqa_plan_size(void* plan) {
  return plan ? static_cast<int>(static_cast<Plan*>(plan)->ops.size()) : 0;
}

static int replay(Plan* p, cudaStream_t s) {
  for (auto& op : p->ops) {
    int r = op(s);
    if (r) return r;
  }
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

}  // extern "C"
