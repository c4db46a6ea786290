// pool_probe.cu -- the allocation pattern of one create / anneal (random-site, shared initial
// fields) / destroy cycle of the library against a private stream-ordered pool: time of every
// allocation and the pool's reserved bytes per cycle.  ORDER=0: the library's order (upload buffer
// first); ORDER=1: upload buffer last; ORDER=2: fields reserved at create.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
  const int order = argc > 1 ? atoi(argv[1]) : 0;
  cudaMemPoolProps props = {};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = 0;
  cudaMemPool_t pool;
  CK(cudaSetDevice(0));
  CK(cudaMemPoolCreate(&pool, &props));
  unsigned long long thr = ~0ull;
  CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  const size_t A = 134217728, B = 67108864, C = 134217728, F = 268435456, S = 8388608;
  void *host = nullptr;
  CK(cudaMallocHost(&host, A));
  for (int cycle = 0; cycle < 8; ++cycle) {
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    std::vector<void *> live;
    double worst = 0; const char *worst_name = "";
    auto alloc = [&](const char *name, size_t bytes, void **out) {
      const double t0 = now();
      cudaError_t e = cudaMallocFromPoolAsync(out, bytes, pool, st);
      const double dt = now() - t0;
      if (dt > worst) worst = dt, worst_name = name;
      return e;
    };
    const double t0 = now();
    void *a = nullptr, *bad, *b, *dg, *c, *f = nullptr, *br, *en, *ss, *ts;
    if (order != 1) CK(alloc("upload", A, &a));
    CK(alloc("bad", 4, &bad));
    CK(alloc("qoff", B, &b));
    CK(alloc("diag", 16384, &dg));
    CK(alloc("q64", C, &c));
    if (order == 1) CK(alloc("upload", A, &a));
    if (order == 2) CK(alloc("fields", F, &f));
    CK(cudaMemcpyAsync(a, host, A, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(b, 0, B, st));
    CK(cudaMemsetAsync(c, 0, C, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFreeAsync(a, st));
    CK(cudaFreeAsync(bad, st));
    const double t1 = now();
    CK(alloc("best_rel", 131072, &br));
    CK(alloc("energy", 131072, &en));
    CK(alloc("states", S, &ss));
    CK(alloc("tscale", 32768, &ts));
    if (order != 2) CK(alloc("fields", F, &f));
    CK(cudaMemsetAsync(f, 0, F, st));
    CK(cudaStreamSynchronize(st));
    const double t2 = now();
    for (void *q : {b, dg, c, br, en, ss, ts, f}) CK(cudaFreeAsync(q, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamDestroy(st));
    const double t3 = now();
    unsigned long long reserved = 0, used = 0;
    CK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
    CK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used));
    printf("order %d cycle %d: create %.1f ms anneal-allocs %.1f ms destroy %.1f ms; slowest alloc %s %.1f ms; reserved %.0f MiB used %.0f MiB\n",
           order, cycle, t1 - t0, t2 - t1, t3 - t2, worst_name, worst, reserved / 1048576.0, used / 1048576.0);
  }
  return 0;
}
