// p2v_engine.cu - whole-model integer forward of a P2-ViT DeiT/ViT behind the C ABI.
//
// Replaces VisionTransformer.forward in quantized mode (reference: models/vit_fquant.py:700-799): the
// engine owns nothing but a copy of the plan descriptor; weights, per-channel vectors, the workspace and
// the I/O buffers belong to the caller.  One forward is a fixed sequence of kernels
//   quant+patchify, patch-embed GEMM, token assembly,
//   depth x { LN1, qkv GEMM, attention, proj GEMM(+residual), LN2, fc1 GEMM(+GELU), fc2 GEMM(+residual) },
//   final LN (CLS rows), head GEMM
// on one stream, optionally captured once into a CUDA graph and replayed.
#include <cuda.h>
#include <stdarg.h>

#include <string>
#include <vector>

#include "p2v_common.cuh"

namespace p2v {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static int g_pdl = 7;
bool pdl_enabled(int kind) { return (g_pdl & kind) != 0; }

int make_tmap_kmajor(CUtensorMap* map, const void* ptr, int64_t rows, int64_t k, int64_t ld);
int gemm_configure();
int attention_configure_once();
int gemm_i8_tc(const CUtensorMap& ta, const CUtensorMap& tb, int8_t* out, int64_t ld_out, int m, int n, int k,
               const p2v_epilogue& epi, cudaStream_t st);

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct DumpEntry {
  std::string name;
  int64_t offset, bytes;
  int32_t elem;
};

}  // namespace p2v

using namespace p2v;

// Makes `device` current for the lifetime of the object and restores the caller's device afterwards: the library must
// not change the device a torch program believes to be current.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

struct p2v_vit {
  p2v_vit_desc d;
  std::vector<p2v_block_desc> blocks;
  int device = 0;
  int np = 0, ntok = 0, k0 = 0;
  // weight tensor maps (static)
  CUtensorMap tm_w_pe, tm_w_head;
  std::vector<CUtensorMap> tm_w;  // 4 per block: qkv, proj, fc1, fc2
  // binding to (batch, workspace)
  int bound_b = 0;
  void* bound_ws = nullptr;
  int8_t *patches = nullptr, *pe = nullptr, *x0 = nullptr, *x1 = nullptr, *a = nullptr, *qkv = nullptr,
         *o = nullptr, *hid = nullptr, *cls = nullptr, *logit_codes = nullptr;
  CUtensorMap tm_patches, tm_a_d, tm_o, tm_hid, tm_cls;
  // graph cache: one captured launch sequence per (x, logits, codes, batch, workspace) binding, so that
  // double-buffered inputs (copy of batch i+1 overlapping the forward of batch i) replay without re-capture
  struct GraphEntry {
    cudaGraphExec_t exec;
    const void* x;
    float* logits;
    int8_t* codes;
    int b;
    void* ws;
  };
  std::vector<GraphEntry> graphs;
  // dump layout cache
  int dump_b = 0;
  std::vector<DumpEntry> dump;
  int64_t dump_total = 0;
};

static void build_dump_layout(p2v_vit* h, int b) {
  if (h->dump_b == b) return;
  h->dump.clear();
  int64_t off = 0;
  const int64_t M = (int64_t)b * h->ntok, Mp = (int64_t)b * h->np, D = h->d.embed_dim;
  auto add = [&](const std::string& name, int64_t elems, int32_t elem) {
    h->dump.push_back({name, off, elems * elem, elem});
    off = align_up(off + elems * elem, 256);
  };
  add("patches", Mp * h->k0, 1);
  add("act/patch_embed.qact", Mp * D, 1);
  add("act/qact1", M * D, 1);
  for (int i = 0; i < h->d.depth; ++i) {
    const std::string p = "blocks." + std::to_string(i);
    add("ln/" + p + ".norm1", M * D, 4);
    add("act/" + p + ".attn.qact0", M * D, 1);
    add("act/" + p + ".attn.qact1", M * 3 * D, 1);
    add("act/" + p + ".attn.qact_attn1", (int64_t)b * h->d.num_heads * h->ntok * h->ntok, 1);
    add("softmax/" + p + ".attn.log_int_softmax", (int64_t)b * h->d.num_heads * h->ntok * h->ntok, 1);
    add("act/" + p + ".attn.qact2", M * D, 1);
    add("act/" + p + ".attn.qact3", M * D, 1);
    add("act/" + p + ".qact2", M * D, 1);
    add("ln/" + p + ".norm2", M * D, 4);
    add("act/" + p + ".mlp.qact0", M * D, 1);
    add("act/" + p + ".mlp.qact1", M * h->d.hidden_dim, 1);
    add("act/" + p + ".mlp.qact2", M * D, 1);
    add("act/" + p + ".qact4", M * D, 1);
  }
  add("ln/norm", (int64_t)b * D, 4);
  add("act/qact2", (int64_t)b * D, 1);
  add("act/act_out", (int64_t)b * h->d.num_classes, 1);
  h->dump_total = off;
  h->dump_b = b;
}

static int64_t ws_layout(const p2v_vit* h, int b, int64_t* offs) {
  const int64_t M = (int64_t)b * h->ntok, Mp = (int64_t)b * h->np, D = h->d.embed_dim;
  const int64_t sizes[10] = {Mp * h->k0, Mp * D, M * D, M * D, M * D, M * 3 * D, M * D, M * h->d.hidden_dim,
                             (int64_t)b * D, (int64_t)b * h->d.num_classes};
  int64_t off = 0;
  for (int i = 0; i < 10; ++i) {
    if (offs) offs[i] = off;
    off = align_up(off + sizes[i], 1024);
  }
  return off;
}

static int bind(p2v_vit* h, int b, void* ws) {
  if (h->bound_b == b && h->bound_ws == ws) return P2V_OK;
  P2V_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "p2v_vit_forward: workspace must be 1024-byte aligned");
  int64_t offs[10];
  ws_layout(h, b, offs);
  int8_t* base = static_cast<int8_t*>(ws);
  h->patches = base + offs[0];
  h->pe = base + offs[1];
  h->x0 = base + offs[2];
  h->x1 = base + offs[3];
  h->a = base + offs[4];
  h->qkv = base + offs[5];
  h->o = base + offs[6];
  h->hid = base + offs[7];
  h->cls = base + offs[8];
  h->logit_codes = base + offs[9];
  const int64_t M = (int64_t)b * h->ntok, Mp = (int64_t)b * h->np, D = h->d.embed_dim;
  int rc;
  if ((rc = make_tmap_kmajor(&h->tm_patches, h->patches, Mp, h->k0, h->k0))) return rc;
  if ((rc = make_tmap_kmajor(&h->tm_a_d, h->a, M, D, D))) return rc;
  if ((rc = make_tmap_kmajor(&h->tm_o, h->o, M, D, D))) return rc;
  if ((rc = make_tmap_kmajor(&h->tm_hid, h->hid, M, h->d.hidden_dim, h->d.hidden_dim))) return rc;
  if ((rc = make_tmap_kmajor(&h->tm_cls, h->cls, b, D, D))) return rc;
  h->bound_b = b;
  h->bound_ws = ws;
  // Captured graphs of other (batch, workspace) bindings stay valid: a captured launch carries its tensor maps by value
  // (__grid_constant__ kernel parameters) and its buffers by address, and the cache is keyed on (x, logits, codes,
  // batch, workspace).  A validation loop whose last batch is shorter, or two alternating batch sizes, re-binds here
  // without throwing the other binding's graphs away.
  return P2V_OK;
}

#define P2V_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != P2V_OK) return _rc; \
  } while (0)

static int copy_dump(p2v_vit* h, void* dump, int& idx, const void* src, cudaStream_t st) {
  if (dump != nullptr && src != nullptr) {
    const DumpEntry& e = h->dump[idx];
    P2V_CHECK_CUDA(cudaMemcpyAsync(static_cast<char*>(dump) + e.offset, src, e.bytes, cudaMemcpyDeviceToDevice, st));
  }
  ++idx;
  return P2V_OK;
}

// The launch sequence.  With dump != nullptr every intermediate is also written to the dump buffer.
struct U8Input {            // non-null x8: the forward starts from 8-bit pixels
  const uint8_t* x8 = nullptr;
  float mean[4] = {0, 0, 0, 0}, stdv[4] = {1, 1, 1, 1};
};

static int run(p2v_vit* h, const float* x, float* logits, int8_t* logit_codes, int b, void* dump, cudaStream_t st,
               const U8Input* u8 = nullptr) {
  const p2v_vit_desc& d = h->d;
  const int M = b * h->ntok, Mp = b * h->np, D = d.embed_dim;
  char* dp = static_cast<char*>(dump);
  int di = 0;
  auto slot = [&](int i) -> void* { return dp ? dp + h->dump[i].offset : nullptr; };

  if (u8 != nullptr && u8->x8 != nullptr)
    P2V_TRY(p2v_quant_patchify_u8(u8->x8, h->patches, b, d.in_chans, d.img_size, d.img_size, d.patch_size, d.input_scale,
                                  d.input_zp, u8->mean, u8->stdv, st));
  else
    P2V_TRY(p2v_quant_patchify(x, h->patches, b, d.in_chans, d.img_size, d.img_size, d.patch_size, d.input_scale,
                               d.input_zp, st));
  P2V_TRY(copy_dump(h, dump, di, h->patches, st));
  P2V_TRY(gemm_i8_tc(h->tm_patches, h->tm_w_pe, h->pe, D, Mp, D, h->k0, d.patch_embed.epi, st));
  P2V_TRY(copy_dump(h, dump, di, h->pe, st));
  P2V_TRY(p2v_embed_assemble(h->pe, h->x0, b, h->np, D, d.pe_scale, d.pe_zp, d.embed_scale, d.embed_zp, d.cls_value,
                             d.pos_value, d.embed_out_scale, st));
  P2V_TRY(copy_dump(h, dump, di, h->x0, st));

  for (int i = 0; i < d.depth; ++i) {
    const p2v_block_desc& bk = h->blocks[i];
    // norm1 + attn.qact0
    P2V_TRY(p2v_layernorm_int(h->x0, D, h->a, static_cast<int32_t*>(slot(di)), M, D, &bk.norm1, st));
    ++di;
    P2V_TRY(copy_dump(h, dump, di, h->a, st));
    // qkv + attn.qact1
    P2V_TRY(gemm_i8_tc(h->tm_a_d, h->tm_w[4 * i + 0], h->qkv, 3 * D, M, 3 * D, D, bk.qkv.epi, st));
    P2V_TRY(copy_dump(h, dump, di, h->qkv, st));
    // attention: qact_attn1, log-int-softmax, qact2
    p2v_attention at = bk.attn;
    at.dump_scores = static_cast<int8_t*>(slot(di));
    at.dump_softmax = static_cast<uint8_t*>(slot(di + 1));
    di += 2;
    P2V_TRY(p2v_attention_int(h->qkv, h->o, b, h->ntok, d.num_heads, &at, st));
    P2V_TRY(copy_dump(h, dump, di, h->o, st));
    // proj + attn.qact3, residual + Block.qact2
    p2v_epilogue ep = bk.proj.epi;
    ep.flags |= P2V_EPI_RESIDUAL;
    ep.residual = h->x0;
    ep.aux_codes = static_cast<int8_t*>(slot(di));
    ++di;
    P2V_TRY(gemm_i8_tc(h->tm_o, h->tm_w[4 * i + 1], h->x1, D, M, D, D, ep, st));
    P2V_TRY(copy_dump(h, dump, di, h->x1, st));
    // norm2 + mlp.qact0
    P2V_TRY(p2v_layernorm_int(h->x1, D, h->a, static_cast<int32_t*>(slot(di)), M, D, &bk.norm2, st));
    ++di;
    P2V_TRY(copy_dump(h, dump, di, h->a, st));
    // fc1 + GELU + mlp.qact1
    P2V_TRY(gemm_i8_tc(h->tm_a_d, h->tm_w[4 * i + 2], h->hid, d.hidden_dim, M, d.hidden_dim, D, bk.fc1.epi, st));
    P2V_TRY(copy_dump(h, dump, di, h->hid, st));
    // fc2 + mlp.qact2, residual + Block.qact4
    ep = bk.fc2.epi;
    ep.flags |= P2V_EPI_RESIDUAL;
    ep.residual = h->x1;
    ep.aux_codes = static_cast<int8_t*>(slot(di));
    ++di;
    P2V_TRY(gemm_i8_tc(h->tm_hid, h->tm_w[4 * i + 3], h->x0, D, M, D, d.hidden_dim, ep, st));
    P2V_TRY(copy_dump(h, dump, di, h->x0, st));
  }
  // final norm on the CLS rows only + qact2
  P2V_TRY(p2v_layernorm_int(h->x0, (int64_t)h->ntok * D, h->cls, static_cast<int32_t*>(slot(di)), b, D, &d.norm, st));
  ++di;
  P2V_TRY(copy_dump(h, dump, di, h->cls, st));
  // head + act_out
  p2v_epilogue ep = d.head.epi;
  ep.flags |= P2V_EPI_OUT_F32;
  ep.out_f32 = logits;
  int8_t* codes = logit_codes ? logit_codes : h->logit_codes;
  P2V_TRY(gemm_i8_tc(h->tm_cls, h->tm_w_head, codes, d.num_classes, b, d.num_classes, D, ep, st));
  P2V_TRY(copy_dump(h, dump, di, codes, st));
  return P2V_OK;
}

extern "C" const char* p2v_last_error(void) { return g_error; }

extern "C" int p2v_set_pdl(int enabled) {
  g_pdl = enabled;
  return P2V_OK;
}
extern "C" int p2v_version(void) { return 100; }

extern "C" int p2v_check_device(int device) {
  cudaDeviceProp prop;
  P2V_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10 || prop.minor != 0) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200): tcgen05 kind::i8 is absent elsewhere",
              device, prop.major, prop.minor);
    return P2V_ERR_UNSUPPORTED;
  }
  return P2V_OK;
}

extern "C" int p2v_vit_create(const p2v_vit_desc* desc, int device, p2v_vit** out) {
  P2V_REQUIRE(desc && out && desc->blocks, "p2v_vit_create: null descriptor");
  P2V_REQUIRE(desc->depth > 0 && desc->embed_dim % 16 == 0 && desc->embed_dim == desc->num_heads * 64,
              "p2v_vit_create: embed_dim=%d must equal num_heads=%d x 64", desc->embed_dim, desc->num_heads);
  P2V_REQUIRE(desc->img_size % desc->patch_size == 0 && desc->patch_size % 16 == 0,
              "p2v_vit_create: img_size=%d patch_size=%d unsupported", desc->img_size, desc->patch_size);
  int rc = p2v_check_device(device);
  if (rc) return rc;
  DeviceGuard guard(device);
  P2V_REQUIRE(guard.ok, "p2v_vit_create: cannot make device %d current", device);
  if ((rc = gemm_configure())) return rc;
  if ((rc = attention_configure_once())) return rc;
  p2v_vit* h = new p2v_vit();
  h->d = *desc;
  h->blocks.assign(desc->blocks, desc->blocks + desc->depth);
  h->d.blocks = h->blocks.data();
  h->device = device;
  const int grid = desc->img_size / desc->patch_size;
  h->np = grid * grid;
  h->ntok = h->np + 1;
  h->k0 = desc->in_chans * desc->patch_size * desc->patch_size;
  const int D = desc->embed_dim;
  rc = make_tmap_kmajor(&h->tm_w_pe, desc->patch_embed.w, D, h->k0, h->k0);
  if (!rc) rc = make_tmap_kmajor(&h->tm_w_head, desc->head.w, desc->num_classes, D, D);
  h->tm_w.resize(4 * desc->depth);
  for (int i = 0; i < desc->depth && !rc; ++i) {
    const p2v_block_desc& bk = h->blocks[i];
    rc = make_tmap_kmajor(&h->tm_w[4 * i + 0], bk.qkv.w, 3 * D, D, D);
    if (!rc) rc = make_tmap_kmajor(&h->tm_w[4 * i + 1], bk.proj.w, D, D, D);
    if (!rc) rc = make_tmap_kmajor(&h->tm_w[4 * i + 2], bk.fc1.w, desc->hidden_dim, D, D);
    if (!rc) rc = make_tmap_kmajor(&h->tm_w[4 * i + 3], bk.fc2.w, D, desc->hidden_dim, desc->hidden_dim);
  }
  if (rc) {
    delete h;
    return rc;
  }
  *out = h;
  return P2V_OK;
}

extern "C" void p2v_vit_destroy(p2v_vit* h) {
  if (!h) return;
  for (auto& e : h->graphs) cudaGraphExecDestroy(e.exec);
  delete h;
}

extern "C" int64_t p2v_vit_workspace_bytes(const p2v_vit* h, int b) { return h && b > 0 ? ws_layout(h, b, nullptr) : 0; }

extern "C" int p2v_vit_launches_per_forward(const p2v_vit* h) { return h ? 3 + 7 * h->d.depth + 2 : 0; }

extern "C" int64_t p2v_vit_dump_bytes(const p2v_vit* h, int b) {
  if (!h || b <= 0) return 0;
  build_dump_layout(const_cast<p2v_vit*>(h), b);
  return h->dump_total;
}

extern "C" int p2v_vit_dump_layout(const p2v_vit* hc, int b, int i, const char** name, int64_t* offset, int64_t* bytes,
                                   int32_t* elem_size) {
  p2v_vit* h = const_cast<p2v_vit*>(hc);
  if (!h || b <= 0) return 0;
  build_dump_layout(h, b);
  if (i >= 0 && i < (int)h->dump.size()) {
    if (name) *name = h->dump[i].name.c_str();
    if (offset) *offset = h->dump[i].offset;
    if (bytes) *bytes = h->dump[i].bytes;
    if (elem_size) *elem_size = h->dump[i].elem;
  }
  return (int)h->dump.size();
}

static int forward_impl(p2v_vit* h, const float* x, const U8Input* u8, float* logits, int8_t* logit_codes, int b,
                        void* workspace, void* dump, int use_graph, void* stream) {
  const void* xkey = u8 ? static_cast<const void*>(u8->x8) : static_cast<const void*>(x);
  P2V_REQUIRE(h && xkey && logits && workspace, "p2v_vit_forward: null pointer");
  P2V_REQUIRE(b > 0, "p2v_vit_forward: batch must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard(h->device);      // the handle's device, whatever is current in the calling thread
  P2V_REQUIRE(guard.ok, "p2v_vit_forward: cannot make device %d current", h->device);
  P2V_TRY(bind(h, b, workspace));
  if (dump != nullptr) {
    build_dump_layout(h, b);
    return run(h, x, logits, logit_codes, b, dump, st, u8);
  }
  if (!use_graph || st == nullptr) return run(h, x, logits, logit_codes, b, nullptr, st, u8);
  cudaGraphExec_t exec = nullptr;
  for (auto& e : h->graphs)
    if (e.x == xkey && e.logits == logits && e.codes == logit_codes && e.b == b && e.ws == workspace) exec = e.exec;
  if (exec == nullptr) {
    cudaGraph_t graph = nullptr;
    P2V_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = run(h, x, logits, logit_codes, b, nullptr, st, u8);
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != P2V_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (ce != cudaSuccess) {
      set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
      return P2V_ERR_CUDA;
    }
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
      set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
      return P2V_ERR_CUDA;
    }
    if (h->graphs.size() >= 16) {
      cudaGraphExecDestroy(h->graphs.front().exec);
      h->graphs.erase(h->graphs.begin());
    }
    h->graphs.push_back({exec, xkey, logits, logit_codes, b, workspace});
  }
  P2V_CHECK_CUDA(cudaGraphLaunch(exec, st));
  return P2V_OK;
}

extern "C" int p2v_vit_forward(p2v_vit* h, const float* x, float* logits, int8_t* logit_codes, int b, void* workspace,
                               void* dump, int use_graph, void* stream) {
  return forward_impl(h, x, nullptr, logits, logit_codes, b, workspace, dump, use_graph, stream);
}

extern "C" int p2v_vit_forward_u8(p2v_vit* h, const uint8_t* x, const float* mean, const float* stdv, float* logits,
                                  int8_t* logit_codes, int b, void* workspace, int use_graph, void* stream) {
  P2V_REQUIRE(h && x && mean && stdv, "p2v_vit_forward_u8: null pointer");
  P2V_REQUIRE(h->d.in_chans <= 4, "p2v_vit_forward_u8: at most 4 input channels");
  U8Input u8;
  u8.x8 = x;
  for (int i = 0; i < h->d.in_chans; ++i) {
    u8.mean[i] = mean[i];
    u8.stdv[i] = stdv[i];
  }
  return forward_impl(h, nullptr, &u8, logits, logit_codes, b, workspace, nullptr, use_graph, stream);
}

extern "C" int p2v_vit_forward_host(p2v_vit* h, const float* x_host, float* logits_host, int b, void* workspace,
                                    void* x_dev, void* logits_dev, void* stream) {
  P2V_REQUIRE(h && x_host && logits_host && x_dev && logits_dev, "p2v_vit_forward_host: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t xin = (size_t)b * h->d.in_chans * h->d.img_size * h->d.img_size * sizeof(float);
  const size_t xout = (size_t)b * h->d.num_classes * sizeof(float);
  P2V_CHECK_CUDA(cudaMemcpyAsync(x_dev, x_host, xin, cudaMemcpyHostToDevice, st));
  P2V_TRY(p2v_vit_forward(h, static_cast<const float*>(x_dev), static_cast<float*>(logits_dev), nullptr, b, workspace,
                          nullptr, 1, stream));
  P2V_CHECK_CUDA(cudaMemcpyAsync(logits_host, logits_dev, xout, cudaMemcpyDeviceToHost, st));
  P2V_CHECK_CUDA(cudaStreamSynchronize(st));
  return P2V_OK;
}
