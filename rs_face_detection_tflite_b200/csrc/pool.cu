// pool.cu -- fdl_pool: every GPU of a box behind one handle (include/fdl.h; SURVEY.md 8e "one host thread + pinned staging per GPU").
// One fdl_pipeline and one worker thread per listed device; the application thread only queues jobs.  No collective: the path has no
// exchange step.
#include <cuda_runtime.h>
#include <sched.h>

#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "device_util.h"
#include "fdl_status.h"

using namespace fdl;

namespace {

struct Job {
  int ticket = -1;
  bool jpeg = false;
  std::vector<fdl_image> frames;
  std::vector<const uint8_t*> data;
  std::vector<size_t> len;
  int n = 0;
  // filled by the worker
  bool submitted = false;
  int rc = FDL_OK;
  std::string err;
  int local_ticket = -1;
};

struct Worker {
  fdl_pipeline* pipe = nullptr;
  int device = 0;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Job*> queue;       // jobs waiting for fdl_pipeline_submit
  bool stop = false;
  int in_flight = 0;            // queued or submitted, not yet collected
};

// Run the calling thread on the CPU cores that are local to `device` (the pinned staging buffers the pipeline allocates from this
// thread are first touched there).  Best effort: sysfs knows the GPU's NUMA node and that node's CPU list.
void pin_thread_near(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return; }
  for (char* p = bus; *p; ++p) if (*p >= 'A' && *p <= 'Z') *p = (char)(*p - 'A' + 'a');
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE* f = fopen(path, "r");
  if (!f) return;
  char list[4096] = {0};
  const size_t got = fread(list, 1, sizeof list - 1, f);
  fclose(f);
  if (!got) return;
  cpu_set_t set;
  CPU_ZERO(&set);
  int any = 0;
  for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &set); ++any; } }
    else if (sscanf(tok, "%d", &a) == 1 && a < CPU_SETSIZE) { CPU_SET(a, &set); ++any; }
  }
  if (any) sched_setaffinity(0, sizeof set, &set);
}

}  // namespace

struct fdl_pool {
  std::vector<Worker*> workers;
  std::mutex mu;                               // tickets
  std::vector<std::pair<Job*, int>> tickets;   // index = ticket: (job, worker) until collected
  int depth_per_device = 4;
  int rr = 0;
};

static void worker_main(Worker* w) {
  pin_thread_near(w->device);
  for (;;) {
    Job* j = nullptr;
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return w->stop || !w->queue.empty(); });
      if (w->queue.empty()) return;             // stop requested and nothing left
      j = w->queue.front();
      w->queue.pop_front();
    }
    int lt = -1, rc;
    if (j->jpeg) rc = fdl_pipeline_submit_jpeg(w->pipe, j->data.data(), j->len.data(), j->n, &lt);
    else rc = fdl_pipeline_submit(w->pipe, j->frames.data(), j->n, &lt);
    {
      std::lock_guard<std::mutex> lk(w->mu);
      j->rc = rc; j->local_ticket = lt; j->submitted = true;
      if (rc) j->err = fdl_last_error();
    }
    w->cv.notify_all();
  }
}

extern "C" {

int fdl_pool_create(const fdl_pipeline_config* cfg, const int* devices, int n_devices, fdl_pool** out) try {
  if (!cfg || !out || !devices || n_devices <= 0) return set_error(FDL_ERR_INVALID, "null argument");
  *out = nullptr;
  fdl_pool* p = new fdl_pool();
  for (int i = 0; i < n_devices; ++i) {
    fdl_pipeline_config c = *cfg;
    c.device = devices[i];
    fdl_pipeline* pipe = nullptr;
    const int rc = fdl_pipeline_create(&c, &pipe);
    if (rc) { const std::string msg = fdl_last_error(); fdl_pool_destroy(p); return set_error(rc, "device " + std::to_string(devices[i]) + ": " + msg); }
    Worker* w = new Worker();
    w->pipe = pipe; w->device = devices[i];
    p->workers.push_back(w);
    p->depth_per_device = fdl_pipeline_depth(pipe);
  }
  for (Worker* w : p->workers) w->th = std::thread(worker_main, w);
  *out = p;
  return FDL_OK;
} FDL_ABI_CATCH

void fdl_pool_destroy(fdl_pool* p) {
  if (!p) return;
  for (Worker* w : p->workers) {
    { std::lock_guard<std::mutex> lk(w->mu); w->stop = true; }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
    fdl_pipeline_destroy(w->pipe);
    delete w;
  }
  for (auto& t : p->tickets) delete t.first;
  delete p;
}

int fdl_pool_devices(const fdl_pool* p) { return p ? (int)p->workers.size() : 0; }
int fdl_pool_depth(const fdl_pool* p) { return p ? (int)p->workers.size() * p->depth_per_device : 0; }

static int pool_enqueue(fdl_pool* p, Job* j, int* ticket) {
  // least-loaded pipeline, round-robin among equals
  const int nw = (int)p->workers.size();
  int best = -1, best_load = 1 << 30;
  for (int k = 0; k < nw; ++k) {
    const int i = (p->rr + k) % nw;
    Worker* w = p->workers[(size_t)i];
    std::lock_guard<std::mutex> lk(w->mu);
    if (w->in_flight < best_load) { best_load = w->in_flight; best = i; }
  }
  if (best_load >= p->depth_per_device) { delete j; return set_error(FDL_ERR_INVALID, "every pipeline of the pool is full: collect a ticket first"); }
  p->rr = (best + 1) % nw;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    j->ticket = (int)p->tickets.size();
    p->tickets.emplace_back(j, best);
  }
  Worker* w = p->workers[(size_t)best];
  { std::lock_guard<std::mutex> lk(w->mu); w->queue.push_back(j); ++w->in_flight; }
  w->cv.notify_all();
  *ticket = j->ticket;
  return FDL_OK;
}

int fdl_pool_submit(fdl_pool* p, const fdl_image* frames, int n, int* ticket) try {
  if (!p || !frames || !ticket || n <= 0) return set_error(FDL_ERR_INVALID, "bad arguments");
  Job* j = new Job();
  j->n = n; j->frames.assign(frames, frames + n);
  return pool_enqueue(p, j, ticket);
} FDL_ABI_CATCH

int fdl_pool_submit_jpeg(fdl_pool* p, const uint8_t* const* data, const size_t* len, int n, int* ticket) try {
  if (!p || !data || !len || !ticket || n <= 0) return set_error(FDL_ERR_INVALID, "bad arguments");
  Job* j = new Job();
  j->jpeg = true; j->n = n; j->data.assign(data, data + n); j->len.assign(len, len + n);
  return pool_enqueue(p, j, ticket);
} FDL_ABI_CATCH

int fdl_pool_collect(fdl_pool* p, int ticket, fdl_frame_result* frame_results, fdl_face_result* face_results, int* n, int* device_index) try {
  if (!p) return set_error(FDL_ERR_INVALID, "null argument");
  Job* j = nullptr;
  int wi = -1;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    if (ticket < 0 || ticket >= (int)p->tickets.size() || !p->tickets[(size_t)ticket].first) return set_error(FDL_ERR_INVALID, "unknown ticket");
    j = p->tickets[(size_t)ticket].first; wi = p->tickets[(size_t)ticket].second;
    p->tickets[(size_t)ticket].first = nullptr;
  }
  Worker* w = p->workers[(size_t)wi];
  {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return j->submitted; });
  }
  int rc = j->rc;
  if (rc == FDL_OK) rc = fdl_pipeline_collect(w->pipe, j->local_ticket, frame_results, face_results, n);   // a lane's results are private to its ticket
  else set_error(rc, j->err);
  if (device_index) *device_index = wi;
  { std::lock_guard<std::mutex> lk(w->mu); --w->in_flight; }
  delete j;
  return rc;
} FDL_ABI_CATCH

}  // extern "C"
