// Host <-> device copy ceiling of one box at 1..N GPUs (no kernels): what the host-buffer leg of bench.py (`e2e`) can
// reach at best.  Per GPU: pinned host buffers, a 36.8 MB device->host chunk (one 12 MP BGR8 frame) and a 12.3 MB
// host->device chunk (one Bayer frame) in flight on separate streams, for a fixed time.  Modes: `threads` (one process,
// one host thread per GPU -- what rip_apply_batch_host_multi does) and `procs` (one process per GPU -- what torchrun does).
//
//   nvcc -O2 -o tools/pcie_probe/probe tools/pcie_probe/probe.cu
//   tools/pcie_probe/probe --gpus 8 --mode threads --seconds 2 [--dir d2h|h2d|both]
// Prints one JSON line.
#include <cuda_runtime.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

struct Result { double d2h_bytes = 0, h2d_bytes = 0, seconds = 0; int ok = 0; };

static Result run_gpu(int dev, double t_start, double seconds, bool do_d2h, bool do_h2d) {
  Result r;
  const size_t out_chunk = (size_t)4032 * 3040 * 3, in_chunk = (size_t)4032 * 3040;
  const int depth = 4;
  if (cudaSetDevice(dev) != cudaSuccess) return r;
  uint8_t *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
  if (cudaMallocHost(&h_in, in_chunk * depth) || cudaMallocHost(&h_out, out_chunk * depth) || cudaMalloc(&d_in, in_chunk * depth) ||
      cudaMalloc(&d_out, out_chunk * depth)) return r;
  memset(h_in, 1, in_chunk * depth);
  cudaMemset(d_out, 2, out_chunk * depth);
  cudaStream_t s_up, s_down;
  cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking);
  cudaDeviceSynchronize();
  while (now() < t_start) usleep(200);
  const double t0 = now();
  long n_up = 0, n_down = 0;
  while (now() - t0 < seconds) {
    for (int i = 0; i < depth; ++i) {
      if (do_h2d) { cudaMemcpyAsync(d_in + i * in_chunk, h_in + i * in_chunk, in_chunk, cudaMemcpyHostToDevice, s_up); ++n_up; }
      if (do_d2h) { cudaMemcpyAsync(h_out + i * out_chunk, d_out + i * out_chunk, out_chunk, cudaMemcpyDeviceToHost, s_down); ++n_down; }
    }
    cudaStreamSynchronize(s_up);
    cudaStreamSynchronize(s_down);
  }
  r.seconds = now() - t0;
  r.d2h_bytes = (double)n_down * out_chunk; r.h2d_bytes = (double)n_up * in_chunk;
  r.ok = cudaGetLastError() == cudaSuccess;
  cudaFreeHost(h_in); cudaFreeHost(h_out); cudaFree(d_in); cudaFree(d_out);
  return r;
}

int main(int argc, char** argv) {
  int gpus = 1; double seconds = 2.0; std::string mode = "threads", dir = "both";
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "--gpus")) gpus = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "--seconds")) seconds = atof(argv[i + 1]);
    else if (!strcmp(argv[i], "--mode")) mode = argv[i + 1];
    else if (!strcmp(argv[i], "--dir")) dir = argv[i + 1];
  }
  const bool d2h = dir != "h2d", h2d = dir != "d2h";
  std::vector<Result> res(gpus);
  const double t_start = now() + 3.0;  // everyone allocates first, then starts together
  if (mode == "threads") {
    int n = 0; cudaGetDeviceCount(&n);
    if (n < gpus) { printf("{\"error\": \"only %d GPUs\"}\n", n); return 1; }
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; ++g) th.emplace_back([&, g] { res[g] = run_gpu(g, t_start, seconds, d2h, h2d); });
    for (auto& t : th) t.join();
  } else {
    std::vector<int> fds(gpus);
    for (int g = 0; g < gpus; ++g) {
      int p[2]; if (pipe(p)) return 1;
      if (fork() == 0) {  // child: CUDA is initialised after the fork
        close(p[0]);
        Result r = run_gpu(g, t_start, seconds, d2h, h2d);
        if (write(p[1], &r, sizeof r) != (ssize_t)sizeof r) _exit(2);
        _exit(0);
      }
      close(p[1]); fds[g] = p[0];
    }
    for (int g = 0; g < gpus; ++g) { if (read(fds[g], &res[g], sizeof(Result)) != (ssize_t)sizeof(Result)) res[g].ok = 0; close(fds[g]); }
    while (wait(nullptr) > 0) {}
  }
  double d2h_gbs = 0, h2d_gbs = 0; int ok = 1;
  std::string per = "[";
  for (int g = 0; g < gpus; ++g) {
    const double a = res[g].seconds > 0 ? res[g].d2h_bytes / res[g].seconds / 1e9 : 0, b = res[g].seconds > 0 ? res[g].h2d_bytes / res[g].seconds / 1e9 : 0;
    d2h_gbs += a; h2d_gbs += b; ok &= res[g].ok;
    char buf[96]; snprintf(buf, sizeof buf, "%s[%.2f, %.2f]", g ? ", " : "", a, b); per += buf;
  }
  per += "]";
  printf("{\"probe\": \"pcie\", \"gpus\": %d, \"mode\": \"%s\", \"dir\": \"%s\", \"ok\": %d, \"d2h_gbs\": %.2f, \"h2d_gbs\": %.2f, \"total_gbs\": %.2f, "
         "\"per_gpu_d2h_h2d_gbs\": %s, \"chunk_bytes\": [36771840, 12257280], \"seconds\": %.2f}\n",
         gpus, mode.c_str(), dir.c_str(), ok, d2h_gbs, h2d_gbs, d2h_gbs + h2d_gbs, per.c_str(), seconds);
  return ok ? 0 : 1;
}
