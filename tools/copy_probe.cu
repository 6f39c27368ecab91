// copy_probe.cu -- what the box's host<->device copy path can do, independent of any kernel.
//   nvcc -O2 -o tools/copy_probe tools/copy_probe.cu
//   tools/copy_probe <device> [h2d_MB d2h_MB [reps]]
// Prints one JSON object: host topology (NUMA nodes, the GPU's node), and GB/s for H2D alone, D2H alone and both at
// once, for pinned buffers allocated (a) by cudaMallocHost from the calling thread as-is, (b) after binding the thread to
// each NUMA node's CPUs (first touch there), (c) write-combined.  Used to set bench.py's e2e copy roofline.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <chrono>

static std::string slurp(const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) return "";
  char buf[4096];
  size_t n = fread(buf, 1, sizeof buf - 1, f);
  fclose(f);
  buf[n] = 0;
  while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ')) buf[--n] = 0;
  return buf;
}

static std::vector<int> parse_cpulist(const std::string &s) {
  std::vector<int> v;
  const char *p = s.c_str();
  while (*p) {
    int a = strtol(p, (char **)&p, 10), b = a;
    if (*p == '-') b = strtol(p + 1, (char **)&p, 10);
    for (int i = a; i <= b; i++) v.push_back(i);
    if (*p == ',') p++;
    else break;
  }
  return v;
}

static bool bind_cpus(const std::vector<int> &cpus) {
  cpu_set_t set;
  CPU_ZERO(&set);
  for (int c : cpus) CPU_SET(c, &set);
  return sched_setaffinity(0, sizeof set, &set) == 0;
}

struct Res { double h2d, d2h, both_h2d, both_d2h; };

static Res measure(void *hin, void *hout, void *din, void *dout, size_t nin, size_t nout, int reps, size_t seg) {
  cudaStream_t s0, s1;
  cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
  auto run = [&](bool in, bool out) {
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) {
      if (in) for (size_t o = 0; o < nin; o += seg)
        cudaMemcpyAsync((char *)din + o, (char *)hin + o, (nin - o < seg) ? nin - o : seg, cudaMemcpyHostToDevice, s0);
      if (out) for (size_t o = 0; o < nout; o += seg)
        cudaMemcpyAsync((char *)hout + o, (char *)dout + o, (nout - o < seg) ? nout - o : seg, cudaMemcpyDeviceToHost, s1);
    }
    cudaDeviceSynchronize();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
  };
  run(true, true);
  Res r;
  r.h2d = nin / run(true, false) / 1e9;
  r.d2h = nout / run(false, true) / 1e9;
  double t = run(true, true);
  r.both_h2d = nin / t / 1e9;
  r.both_d2h = nout / t / 1e9;
  cudaStreamDestroy(s0);
  cudaStreamDestroy(s1);
  return r;
}

int main(int argc, char **argv) {
  int dev = argc > 1 ? atoi(argv[1]) : 0;
  size_t nin = (size_t)(argc > 2 ? atoi(argv[2]) : 1478) * 1000000, nout = (size_t)(argc > 3 ? atoi(argv[3]) : 487) * 1000000;
  int reps = argc > 4 ? atoi(argv[4]) : 5;
  if (cudaSetDevice(dev) != cudaSuccess) { printf("{\"error\": \"no device %d\"}\n", dev); return 1; }
  char bus[64] = {0};
  cudaDeviceGetPCIBusId(bus, sizeof bus, dev);
  for (char *p = bus; *p; p++) *p = tolower(*p);
  std::string node_of_gpu = slurp((std::string("/sys/bus/pci/devices/") + bus + "/numa_node").c_str());
  std::vector<std::vector<int>> nodes;
  for (int n = 0; n < 16; n++) {
    std::string s = slurp(("/sys/devices/system/node/node" + std::to_string(n) + "/cpulist").c_str());
    if (s.empty()) break;
    nodes.push_back(parse_cpulist(s));
  }
  cpu_set_t cur;
  sched_getaffinity(0, sizeof cur, &cur);
  void *din, *dout;
  cudaMalloc(&din, nin);
  cudaMalloc(&dout, nout);
  printf("{\"device\": %d, \"pci\": \"%s\", \"gpu_numa_node\": \"%s\", \"numa_nodes\": %zu, \"allowed_cpus\": %d, "
         "\"h2d_bytes\": %zu, \"d2h_bytes\": %zu, \"runs\": [", dev, bus, node_of_gpu.c_str(), nodes.size(), CPU_COUNT(&cur), nin, nout);
  bool first = true;
  auto report = [&](const char *what, int node, Res r, size_t seg) {
    printf("%s\n {\"alloc\": \"%s\", \"node\": %d, \"segment_mb\": %.1f, \"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, "
           "\"both_h2d_gbs\": %.2f, \"both_d2h_gbs\": %.2f}", first ? "" : ",", what, node, seg / 1e6, r.h2d, r.d2h, r.both_h2d, r.both_d2h);
    first = false;
    fflush(stdout);
  };
  const size_t segs[2] = {27648000, nin};
  // (a) as the process finds itself
  {
    void *hin, *hout;
    cudaMallocHost(&hin, nin); cudaMallocHost(&hout, nout);
    memset(hin, 1, nin); memset(hout, 0, nout);
    for (size_t seg : segs) report("cudaMallocHost", -1, measure(hin, hout, din, dout, nin, nout, reps, seg), seg);
    cudaFreeHost(hin); cudaFreeHost(hout);
  }
  // (b) thread bound to each NUMA node, pages first-touched there, then registered
  for (size_t n = 0; n < nodes.size() && nodes.size() > 1; n++) {
    if (!bind_cpus(nodes[n])) continue;
    void *hin = aligned_alloc(2 << 20, (nin + (2 << 20)) & ~((size_t)(2 << 20) - 1));
    void *hout = aligned_alloc(2 << 20, (nout + (2 << 20)) & ~((size_t)(2 << 20) - 1));
    memset(hin, 1, nin); memset(hout, 0, nout);
    cudaHostRegister(hin, nin, cudaHostRegisterDefault);
    cudaHostRegister(hout, nout, cudaHostRegisterDefault);
    report("first-touch+cudaHostRegister", (int)n, measure(hin, hout, din, dout, nin, nout, reps, segs[0]), segs[0]);
    cudaHostUnregister(hin); cudaHostUnregister(hout);
    free(hin); free(hout);
  }
  sched_setaffinity(0, sizeof cur, &cur);
  // (c) write-combined source
  {
    void *hin, *hout;
    cudaHostAlloc(&hin, nin, cudaHostAllocWriteCombined); cudaMallocHost(&hout, nout);
    memset(hin, 1, nin); memset(hout, 0, nout);
    report("write-combined source", -1, measure(hin, hout, din, dout, nin, nout, reps, segs[0]), segs[0]);
    cudaFreeHost(hin); cudaFreeHost(hout);
  }
  printf("\n]}\n");
  return 0;
}
