// context.cu -- cb2_ctx lifecycle, error reporting, device scratch arena, pinned staging.
#include "common.cuh"

static std::string g_create_error;
static std::mutex g_create_mu;

int cb2_fail(cb2_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else {
    std::lock_guard<std::mutex> g(g_create_mu);
    g_create_error = buf;
  }
  return code;
}

extern "C" const char* cb2_last_error(const cb2_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  return g_create_error.c_str();
}

extern "C" int cb2_create(int device, int precision, cb2_ctx** out) {
  if (!out) return cb2_fail(nullptr, CB2_ERR_INVALID, "cb2_create: out is NULL");
  *out = nullptr;
  if (precision != CB2_FP32 && precision != CB2_FP64)
    return cb2_fail(nullptr, CB2_ERR_INVALID, "cb2_create: precision must be CB2_FP32 or CB2_FP64");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return cb2_fail(nullptr, CB2_ERR_NO_DEVICE,
                    "cb2_create: no CUDA device (%s); libcaesar_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev)
    return cb2_fail(nullptr, CB2_ERR_INVALID, "cb2_create: device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cb2_fail(nullptr, CB2_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return cb2_fail(nullptr, CB2_ERR_NO_DEVICE,
                    "cb2_create: device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                    prop.minor);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return cb2_fail(nullptr, CB2_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cb2_ctx* ctx = new cb2_ctx();
  ctx->device = device;
  ctx->precision = precision;
  ctx->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return cb2_fail(nullptr, CB2_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  ctx->stream = ctx->own_stream;
  for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->copy_ev[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->small_dev, 64 * sizeof(double));
  if (e != cudaSuccess) {  // unwind whatever was created
    const int code = e == cudaErrorMemoryAllocation ? CB2_ERR_NOMEM : CB2_ERR_CUDA;
    cudaGetLastError();
    for (int i = 0; i < 8; ++i) {
      if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
      if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return cb2_fail(nullptr, code, "cb2_create: events / scratch: %s", cudaGetErrorString(e));
  }
  *out = ctx;
  return CB2_OK;
}

extern "C" void cb2_destroy(cb2_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->pending) {
    if (kv.second.done) cudaEventDestroy(kv.second.done);
    if (kv.second.dev_block) cudaFree(kv.second.dev_block);
  }
  for (int i = 0; i < 8; ++i) {
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
  }
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->arena.base) cudaFree(ctx->arena.base);
  if (ctx->io.base) cudaFree(ctx->io.base);
  if (ctx->small_dev) cudaFree(ctx->small_dev);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" int cb2_precision(const cb2_ctx* ctx) { return ctx ? ctx->precision : CB2_ERR_INVALID; }

extern "C" int cb2_set_stream(cb2_ctx* ctx, void* s) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return CB2_OK;
}

extern "C" int cb2_synchronize(cb2_ctx* ctx) {
  if (!ctx) return CB2_ERR_INVALID;
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CB2_OK;
}

// ---- device-resident beliefs: plain device buffers a caller keeps between calls; the *_dev entry points take
// them as they are.  Copies are stream-ordered on the context stream; only the download waits.
extern "C" int cb2_device_alloc(cb2_ctx* ctx, size_t bytes, void** out) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, out && bytes > 0, "null out or zero bytes");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { cudaGetLastError(); return cb2_fail(ctx, CB2_ERR_NOMEM, "device buffer of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
  *out = p;
  return CB2_OK;
}
extern "C" int cb2_device_free(cb2_ctx* ctx, void* ptr) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  if (!ptr) return CB2_OK;
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // enqueued work may still use it
  CB2_CUDA(ctx, cudaFree(ptr));
  return CB2_OK;
}
extern "C" int cb2_upload(cb2_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, dst_dev && src_host, "null pointer");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  CB2_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return CB2_OK;
}
extern "C" int cb2_download(cb2_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, dst_host && src_dev, "null pointer");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  CB2_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CB2_OK;
}

extern "C" uint64_t cb2_kernel_launches(const cb2_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int cb2_device_sm_count(const cb2_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" double cb2_last_stage_ms(const cb2_ctx* ctx, const char* stage) {
  if (!ctx || !stage) return -1.0;
  cb2_lock g(const_cast<cb2_ctx*>(ctx)->mu);  // resolves pending events: mutates the stage table and waits for the stream
  cb2_resolve_product_stages(const_cast<cb2_ctx*>(ctx));
  auto it = ctx->stage_ms.find(stage);
  return it == ctx->stage_ms.end() ? -1.0 : it->second;
}

int cb2_arena_reserve(cb2_ctx* ctx, size_t bytes) {
  bytes = cb2_align(bytes, 1 << 20);
  ctx->arena.off = 0;
  if (bytes <= ctx->arena.cap) return CB2_OK;
  // the old block may still be in use by enqueued work on the stream
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->arena.base) cudaFree(ctx->arena.base);
  ctx->arena.base = nullptr;
  ctx->arena.cap = 0;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cb2_fail(ctx, CB2_ERR_NOMEM, "device scratch of %zu bytes: %s", bytes, cudaGetErrorString(e));
  }
  ctx->arena.base = (char*)p;
  ctx->arena.cap = bytes;
  return CB2_OK;
}

void* cb2_arena_alloc(cb2_ctx* ctx, size_t bytes) {
  size_t o = cb2_align(ctx->arena.off);
  if (o + bytes > ctx->arena.cap) return nullptr;
  ctx->arena.off = o + bytes;
  return ctx->arena.base + o;
}

void* cb2_io_block(cb2_ctx* ctx, size_t bytes) {
  bytes = cb2_align(bytes, 1 << 20);
  if (bytes <= ctx->io.cap) return ctx->io.base;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cb2_fail(ctx, CB2_ERR_CUDA, "stream sync before I/O block growth failed"); return nullptr; }
  if (ctx->io.base) cudaFree(ctx->io.base);
  ctx->io.base = nullptr;
  ctx->io.cap = 0;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cb2_fail(ctx, CB2_ERR_NOMEM, "device I/O block of %zu bytes: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  ctx->io.base = (char*)p;
  ctx->io.cap = bytes;
  return p;
}

int cb2_pinned_reserve(cb2_ctx* ctx, size_t bytes) {
  bytes = cb2_align(bytes, 1 << 16);
  if (bytes <= ctx->pinned_cap) return CB2_OK;
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr;
  ctx->pinned_cap = 0;
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cb2_fail(ctx, CB2_ERR_NOMEM, "pinned staging of %zu bytes: %s", bytes, cudaGetErrorString(e));
  }
  ctx->pinned = (char*)p;
  ctx->pinned_cap = bytes;
  return CB2_OK;
}
