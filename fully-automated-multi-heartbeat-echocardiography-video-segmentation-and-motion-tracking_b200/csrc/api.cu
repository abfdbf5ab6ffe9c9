// api.cu - the C ABI of libclasfv_b200.so (include/clasfv_b200.h): handle lifetime, weight folding and
// packing, the layer schedule of the network, and thin argument-checking wrappers over the kernels.
#include "internal.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>

namespace clasfv {

static thread_local char g_error[1024] = "";

cudaError_t allow_max_dynamic_smem_impl(const void* kernel) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& kd : done) if (kd.first == kernel && kd.second == dev) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) done.emplace_back(kernel, dev);
  return e;
}

bool pdl_enabled() {
  // opt-in: measured on the headline workload (10 steps each way) 20.40 ms with the attribute, 20.20 ms without - consecutive
  // launches of one stream already overlap their launch latency, and a CTA of the next kernel cannot share an SM with one of
  // the current kernel (shared memory), so there is no prologue to hide
  static const bool on = getenv("CLASFV_PDL") != nullptr;
  return on;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

namespace {

constexpr float BN_EPS = 1e-5f;
constexpr int DEC = 64;                 // decoder width
constexpr int STEM_MID = 45, STEM_MID_PAD = 64;
constexpr int DENSE_DEPTH = 5;          // temporal convolutions through the stem and layer1: frames a clip edge reaches

struct HostTensor { std::vector<int64_t> shape; std::vector<float> data; };

struct PackedConv {
  int cin = 0, cout = 0, cin_pad = 0, cout_pad = 0;
  int kt = 1, kh = 1, kw = 1, st = 1, sh = 1, sw = 1, pt = 0, ph = 0, pw = 0;
  void* w = nullptr;          // device [tap][cout_pad][cin_pad], fp32 or bf16
  float* bias = nullptr;      // device [cout_pad] or nullptr
};

struct Block { PackedConv s1, t1, s2, t2, down; bool has_down = false; };

// Small host->device tables (clip offsets, fusion plans) go through a ring of pinned slots so that
// back-to-back asynchronous calls never overwrite a slot whose copy is still in flight.
struct TableRing {
  static constexpr int SLOTS = 8;
  static constexpr size_t SLOT_BYTES = 1 << 20;
  char* host = nullptr; char* dev = nullptr; cudaEvent_t ev[SLOTS]; bool used[SLOTS]; int next = 0;
  int init() {
    CLASFV_CUDA(cudaMallocHost(&host, SLOTS * SLOT_BYTES));
    CLASFV_CUDA(cudaMalloc(&dev, SLOTS * SLOT_BYTES));
    for (int i = 0; i < SLOTS; ++i) { CLASFV_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); used[i] = false; }
    return CLASFV_OK;
  }
  void destroy() {
    if (host) cudaFreeHost(host);
    if (dev) cudaFree(dev);
    if (host) for (int i = 0; i < SLOTS; ++i) cudaEventDestroy(ev[i]);
    host = dev = nullptr;
  }
  // stage `bytes` (already laid out by fill(host_slot)) and return the device address
  template <typename Fill>
  int upload(size_t bytes, cudaStream_t stream, Fill fill, void** dev_ptr) {
    if (bytes > SLOT_BYTES) { set_error("table of %zu bytes exceeds the staging slot", bytes); return CLASFV_EINVAL; }
    const int s = next; next = (next + 1) % SLOTS;
    if (used[s]) CLASFV_CUDA(cudaEventSynchronize(ev[s]));
    fill(host + (size_t)s * SLOT_BYTES);
    CLASFV_CUDA(cudaMemcpyAsync(dev + (size_t)s * SLOT_BYTES, host + (size_t)s * SLOT_BYTES, bytes, cudaMemcpyHostToDevice, stream));
    CLASFV_CUDA(cudaEventRecord(ev[s], stream));
    used[s] = true;
    *dev_ptr = dev + (size_t)s * SLOT_BYTES;
    return CLASFV_OK;
  }
};

}  // namespace
}  // namespace clasfv

using namespace clasfv;

struct clasfv_handle {
  int device = 0, num_sms = 0;
  std::map<std::string, HostTensor> tensors;
  bool finalized = false;
  int precision = CLASFV_F32;
  bool force_simt = false;
  std::vector<void*> dev_allocs;       // packed weights
  // packed network
  float* stem_w = nullptr; float* stem_b = nullptr;
  PackedConv stem_t;
  Block blocks[4][2];
  PackedConv lateral[5];
  float *b1 = nullptr, *w2 = nullptr, *b2 = nullptr, *wh = nullptr, *bh = nullptr;
  std::map<std::pair<int, int>, void*> head_tabs;   // interpolation matrices of the tensor-core head per (H, W)
  // workspace
  void* ws = nullptr; size_t ws_bytes = 0;
  uint32_t* minmax = nullptr;          // per-channel {min, max} bit patterns of clasfv_ingest_u8
  TableRing ring;
  // optional stage profiler: an event after every stage of every (sub-)batch; the time since the previous event
  // of the same call is attributed to that stage.  prof_gflop counts the MACs the convolutions really performed.
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<int> prof_stage;
  size_t prof_used = 0;
  int prof_calls = 0;
  int cur_stage = 0;
  double prof_gflop[4] = {0, 0, 0, 0};
  int64_t launches = 0;        // kernels this handle has launched since it was created (clasfv_launch_count)
  // options (clasfv_set_option)
  int sub_batch = 32;          // clips per internal batch of clasfv_forward
  bool dense_video = true;     // share layer-1 work between overlapping windows of one video (bf16 tensor-core path)
  bool umma_pair = true;       // CTA pairs (cta_group::2) for the convolutions with >= 128 output columns
};

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

const HostTensor* find_tensor(const clasfv_handle* h, const std::string& key) {
  auto it = h->tensors.find(key);
  return it == h->tensors.end() ? nullptr : &it->second;
}

int need(const clasfv_handle* h, const std::string& key, std::vector<int64_t> shape, const HostTensor** out) {
  const HostTensor* t = find_tensor(h, key);
  if (!t) { set_error("finalize: state_dict tensor '%s' was never set", key.c_str()); return CLASFV_ESTATE; }
  if (t->shape != shape) { set_error("finalize: tensor '%s' has the wrong shape", key.c_str()); return CLASFV_EINVAL; }
  *out = t;
  return CLASFV_OK;
}

// BatchNorm (inference) as y = x*scale + shift
int bn_affine(const clasfv_handle* h, const std::string& key, int c, std::vector<float>* scale, std::vector<float>* shift) {
  const HostTensor *g, *b, *m, *v;
  int rc;
  if ((rc = need(h, key + ".weight", {c}, &g))) return rc;
  if ((rc = need(h, key + ".bias", {c}, &b))) return rc;
  if ((rc = need(h, key + ".running_mean", {c}, &m))) return rc;
  if ((rc = need(h, key + ".running_var", {c}, &v))) return rc;
  scale->resize(c); shift->resize(c);
  for (int i = 0; i < c; ++i) {
    const float s = g->data[i] / std::sqrt(v->data[i] + BN_EPS);
    (*scale)[i] = s; (*shift)[i] = b->data[i] - m->data[i] * s;
  }
  return CLASFV_OK;
}

int dev_upload(clasfv_handle* h, const void* src, size_t bytes, void** out) {
  void* d = nullptr;
  CLASFV_CUDA(cudaMalloc(&d, bytes));
  h->dev_allocs.push_back(d);
  CLASFV_CUDA(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
  *out = d;
  return CLASFV_OK;
}

// Pack (Cout,Cin,kt,kh,kw) fp32 -> [tap][cout_pad][cin_pad] with an optional per-output-channel scale,
// zero-filled padding, in fp32 or bf16; upload.
int pack_weight(clasfv_handle* h, const float* w, int cout, int cin, int cin_off, int cin_total, int taps, const float* scale,
                int cout_pad, int cin_pad, int dtype, void** out) {
  const size_t n = (size_t)taps * cout_pad * cin_pad;
  std::vector<float> packed(n, 0.f);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int tap = 0; tap < taps; ++tap) {
        const float v = w[((size_t)co * cin_total + cin_off + ci) * taps + tap] * (scale ? scale[co] : 1.f);
        packed[((size_t)tap * cout_pad + co) * cin_pad + ci] = v;
      }
  if (dtype == CLASFV_F32) return dev_upload(h, packed.data(), n * sizeof(float), out);
  if (dtype == CLASFV_F16) {
    std::vector<__half> ph(n);
    for (size_t i = 0; i < n; ++i) ph[i] = __float2half_rn(packed[i]);
    return dev_upload(h, ph.data(), n * sizeof(__half), out);
  }
  std::vector<__nv_bfloat16> pb(n);
  for (size_t i = 0; i < n; ++i) pb[i] = __float2bfloat16_rn(packed[i]);
  return dev_upload(h, pb.data(), n * sizeof(__nv_bfloat16), out);
}

int pack_conv_bn(clasfv_handle* h, const std::string& conv_key, const std::string& bn_key, int cin, int cout, int kt, int kh, int kw,
                 int st, int sh, int sw, int pt, int ph, int pw, int cin_pad, int cout_pad, PackedConv* pc) {
  const HostTensor* w;
  int rc;
  if ((rc = need(h, conv_key + ".weight", {cout, cin, kt, kh, kw}, &w))) return rc;
  std::vector<float> scale, shift;
  if ((rc = bn_affine(h, bn_key, cout, &scale, &shift))) return rc;
  pc->cin = cin; pc->cout = cout; pc->cin_pad = cin_pad; pc->cout_pad = cout_pad;
  pc->kt = kt; pc->kh = kh; pc->kw = kw; pc->st = st; pc->sh = sh; pc->sw = sw; pc->pt = pt; pc->ph = ph; pc->pw = pw;
  if ((rc = pack_weight(h, w->data.data(), cout, cin, 0, cin, kt * kh * kw, scale.data(), cout_pad, cin_pad, h->precision, &pc->w))) return rc;
  std::vector<float> bias(cout_pad, 0.f);
  for (int i = 0; i < cout; ++i) bias[i] = shift[i];
  return dev_upload(h, bias.data(), bias.size() * sizeof(float), reinterpret_cast<void**>(&pc->bias));
}

void free_packed(clasfv_handle* h) {
  for (void* p : h->dev_allocs) cudaFree(p);
  h->dev_allocs.clear();
  h->finalized = false;
}

inline int midplanes(int inplanes, int planes) { return (inplanes * planes * 27) / (inplanes * 9 + 3 * planes); }

int ensure_workspace(clasfv_handle* h, size_t bytes) {
  if (bytes <= h->ws_bytes) return CLASFV_OK;
  if (h->ws) { CLASFV_CUDA(cudaDeviceSynchronize()); CLASFV_CUDA(cudaFree(h->ws)); h->ws = nullptr; h->ws_bytes = 0; }
  cudaError_t e = cudaMalloc(&h->ws, bytes);
  if (e != cudaSuccess) { set_error("workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return CLASFV_ENOMEM; }
  h->ws_bytes = bytes;
  return CLASFV_OK;
}

ConvArgs make_conv(const PackedConv& pc, int n, int ti, int hi, int wi, const void* in, void* out, const void* residual, int relu,
                   int act_dtype, int out_f32) {
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  ConvShape& s = a.s;
  s.n = n; s.ti = ti; s.hi = hi; s.wi = wi; s.cin = pc.cin_pad; s.cout = pc.cout_pad;
  s.kt = pc.kt; s.kh = pc.kh; s.kw = pc.kw; s.st = pc.st; s.sh = pc.sh; s.sw = pc.sw; s.pt = pc.pt; s.ph = pc.ph; s.pw = pc.pw;
  s.to = (ti + 2 * pc.pt - pc.kt) / pc.st + 1; s.ho = (hi + 2 * pc.ph - pc.kh) / pc.sh + 1; s.wo = (wi + 2 * pc.pw - pc.kw) / pc.sw + 1;
  a.in = in; a.in2 = nullptr; a.in_batch_stride = 0; a.weight = pc.w; a.bias = pc.bias; a.residual = residual; a.out = out;
  a.act_dtype = act_dtype; a.out_f32 = out_f32; a.relu = relu;
  a.macs_per_pos = (double)pc.cin * pc.cout * pc.kt * pc.kh * pc.kw;
  return a;
}

int run_conv(clasfv_handle* h, const ConvArgs& a, cudaStream_t stream) {
  if (h->profiling) h->prof_gflop[h->cur_stage] += 2e-9 * a.macs_per_pos * (double)a.s.n * a.s.to * a.s.ho * a.s.wo;
  // CLASFV_CONV_TRACE=1: time every convolution on its own (stream drained before and after) and print one line per
  // launch to stderr - a development aid for finding the layers furthest from the roofline, never on in a measurement
  static const bool trace = getenv("CLASFV_CONV_TRACE") != nullptr;
  ++h->launches;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (trace) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaStreamSynchronize(stream); cudaEventRecord(e0, stream); }
  ConvArgs ap = a;
  ap.no_pair = h->umma_pair ? 0 : 1;
  const int rc = (a.act_dtype != CLASFV_F32 && !h->force_simt) ? launch_conv_umma(ap, h->num_sms, stream) : launch_conv_simt(a, stream);
  if (trace) {
    cudaEventRecord(e1, stream); cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    const double gf = 2e-9 * a.macs_per_pos * (double)a.s.n * a.s.to * a.s.ho * a.s.wo;
    const ConvShape& s = a.s;
    fprintf(stderr, "conv n=%d in=%dx%dx%dx%d out=%dx%dx%dx%d k=%dx%dx%d s=%d,%d,%d seg=%d res=%d relu=%d  %.4f ms  %.2f GFLOP  %.0f TFLOP/s\n", s.n, s.ti, s.hi,
            s.wi, s.cin, s.to, s.ho, s.wo, s.cout, s.kt, s.kh, s.kw, s.st, s.sh, s.sw, a.seg.on, a.residual ? 1 : 0, a.relu, ms, gf, gf / ms);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  return rc;
}


// interpolation matrices of the tensor-core head for ha's frame geometry: built on first use, kept by the handle
int head_table(clasfv_handle* h, HeadArgs* ha, cudaStream_t stream) {
  auto it = h->head_tabs.find({ha->h, ha->w});
  if (it == h->head_tabs.end()) {
    const size_t bytes = head_table_bytes(*ha);
    CLASFV_REQUIRE(bytes > 0, "the tensor-core head does not tile a %d x %d frame (H %% 8, W %% 16, at most 512 x 512)", ha->h, ha->w);
    void* tab = nullptr;
    CLASFV_CUDA(cudaMalloc(&tab, bytes));
    ++h->launches;
    const int rc = launch_head_table(*ha, tab, stream);
    if (rc) { cudaFree(tab); return rc; }
    it = h->head_tabs.emplace(std::make_pair(ha->h, ha->w), tab).first;
  }
  ha->a_tab = it->second;
  return CLASFV_OK;
}

// ------------------------------------------------------------------------------------------- forward
// One clasfv_forward call.  The clips are processed in internal batches of h->sub_batch.  Two schedules:
//
//  per-clip   every clip runs the whole trunk (any input layout; fp32 mode; short batches).
//
//  dense video (bf16 tensor-core path, clips = equally spaced windows of one resident video)
//             Through the stem and layer1 the network has seen only 5 temporal 3x1x1 convolutions, so frame
//             j of a clip depends on the clip's zero padding only if j < 5 or j > T-6.  All other frames are the
//             same numbers in every window that contains them.  They are computed ONCE over the union of the
//             frames ("video-level" maps V_d / VS_d after the d-th temporal convolution / the spatial convolution
//             before it), and per clip only the d frames next to each clip edge are recomputed at depth d, the
//             temporal convolutions reading a virtual clip spliced from the clip's edge frames and the shared
//             video-level frames (ConvArgs::TimeSeg / ResSeg).  Layer1 is 55 % of the trunk's FLOPs; for stride-1
//             windows this does 21 % of them.  Results are bit-identical to the per-clip schedule (same kernel,
//             same K order per output element).
struct Forward {
  clasfv_handle* h; cudaStream_t stream;
  const float* x; const int64_t* offs_host; int64_t channel_stride;
  int n, t, height, width, out_kind, out_dtype;
  char* seg; char* motion;
  // derived
  int act; size_t es; bool tc_head; size_t gs;
  int T[5], H[5], W[5], C[5];
  char* ws = nullptr;
  size_t o_s0, o_f[5], o_mid, o_ta, o_x1, o_ds, o_g[4], o_gt[4];
  size_t total = 0;
  size_t region(size_t bytes) { size_t o = total; total += (bytes + 255) & ~(size_t)255; return o; }

  int mark(int stage) {
    h->cur_stage = stage < 0 ? 0 : (stage + 1 < 4 ? stage + 1 : 3);
    if (!h->profiling) return CLASFV_OK;
    if (h->prof_events.size() <= h->prof_used) {
      cudaEvent_t ev;
      CLASFV_CUDA(cudaEventCreate(&ev));
      h->prof_events.push_back(ev);
    }
    CLASFV_CUDA(cudaEventRecord(h->prof_events[h->prof_used], stream));
    h->prof_stage.push_back(stage);
    ++h->prof_used;
    return CLASFV_OK;
  }
  void set_stage(int stage) { h->cur_stage = stage; }

  void carve_batch(int nb, bool dense) {
    int64_t P[5];
    for (int i = 0; i < 5; ++i) P[i] = (int64_t)nb * T[i] * H[i] * W[i];
    int64_t mid_elems = 0;
    for (int l = dense ? 1 : 0; l < 4; ++l)
      for (int b = 0; b < 2; ++b) {
        const Block& blk = h->blocks[l][b];
        // s1 output lives at the block input's temporal extent and the block output's spatial extent
        const int64_t e1 = (int64_t)nb * (b == 0 ? T[l] : T[l + 1]) * H[l + 1] * W[l + 1] * blk.s1.cout_pad;
        const int64_t e2 = P[l + 1] * blk.s2.cout_pad;
        mid_elems = std::max(mid_elems, std::max(e1, e2));
      }
    o_s0 = dense ? 0 : region(P[0] * STEM_MID_PAD * es);
    // dense-video schedule: the stem / layer1 outputs of a clip are never assembled (edge frames + the shared video-level map)
    for (int i = 0; i < 5; ++i) o_f[i] = (dense && i < 2) ? 0 : region(P[i] * C[i] * es);
    o_mid = region(mid_elems * es);
    o_ta = region(P[dense ? 2 : 1] * (dense ? 128 : 64) * es);
    o_x1 = region(P[dense ? 2 : 1] * (dense ? 128 : 64) * es);
    o_ds = region(P[2] * 128 * es);
    // level 0 of the dense-video schedule holds only the 2 * DENSE_DEPTH edge frames of every clip
    for (int i = 0; i < 4; ++i) o_g[i] = region((dense && i == 0 ? (int64_t)nb * 2 * DENSE_DEPTH * H[1] * W[1] : P[i + 1]) * DEC * gs);
    // tensor-core head: lateral maps of levels 2-4 interpolated along T to the output's frame count
    for (int i = 1; i < 4; ++i) o_gt[i] = tc_head ? region((int64_t)nb * t * H[i + 1] * W[i + 1] * DEC * gs) : 0;
  }

  // one residual block on a batch of nb clips
  int run_block(const Block& blk, int nb, const void* in, int ti, int hi, int wi, void* out, int* to, int* ho, int* wo,
                const ConvArgs::FrameSel* fsel = nullptr, int64_t in_batch_stride = 0) {
    int rc;
    ConvArgs c1 = make_conv(blk.s1, nb, ti, hi, wi, in, ws + o_mid, nullptr, 1, act, 0);
    if (fsel) { c1.fsel = *fsel; c1.in_batch_stride = in_batch_stride; }
    if ((rc = run_conv(h, c1, stream))) return rc;
    ConvArgs c2 = make_conv(blk.t1, nb, c1.s.to, c1.s.ho, c1.s.wo, ws + o_mid, ws + o_ta, nullptr, 1, act, 0);
    if ((rc = run_conv(h, c2, stream))) return rc;
    ConvArgs c3 = make_conv(blk.s2, nb, c2.s.to, c2.s.ho, c2.s.wo, ws + o_ta, ws + o_mid, nullptr, 1, act, 0);
    if ((rc = run_conv(h, c3, stream))) return rc;
    const void* res = in;
    if (blk.has_down) {
      ConvArgs cd = make_conv(blk.down, nb, ti, hi, wi, in, ws + o_ds, nullptr, 0, act, 0);
      if (fsel) { cd.fsel = *fsel; cd.in_batch_stride = in_batch_stride; }
      if ((rc = run_conv(h, cd, stream))) return rc;
      res = ws + o_ds;
    } else if (fsel) {
      set_error("internal: a frame-selected block input needs a downsample branch (the residual would read the virtual clip)");
      return CLASFV_EINVAL;
    }
    ConvArgs c4 = make_conv(blk.t2, nb, c3.s.to, c3.s.ho, c3.s.wo, ws + o_mid, out, res, 1, act, 0);
    if ((rc = run_conv(h, c4, stream))) return rc;
    *to = c4.s.to; *ho = c4.s.ho; *wo = c4.s.wo;
    return CLASFV_OK;
  }

  // What the dense-video schedule hands to run_tail instead of assembled per-clip stem / layer1 maps
  struct DenseIn {
    const char* e1; const char* e5;       // per clip: the 2 edge frames of the stem output, the 2 * DENSE_DEPTH edge frames of layer1's
    const char* v1; const char* v5;       // video-level maps, at the first frame of this batch's first clip
    const char* g0_video;                 // video-level lateral map of level 0, same origin
    int fs, tv_left;                      // frames between clips; video frames from that origin on
  };

  // layers first_layer..4, lateral projections and the head, for a batch whose f[0] (and f[1] if first_layer == 1) are in place -
  // or, with `di`, are read as virtual clips (edge frames + video-level map)
  int run_tail(int nb, int c0, int first_layer, const DenseIn* di = nullptr) {
    int rc;
    const int64_t F64 = (int64_t)H[0] * W[0] * 64;
    const int D = DENSE_DEPTH;
    for (int l = first_layer; l < 4; ++l) {
      const void* in = ws + o_f[l];
      int ti = T[l], hi = H[l], wi = W[l];
      for (int b = 0; b < 2; ++b) {
        void* out = b == 0 ? (void*)(ws + o_x1) : (void*)(ws + o_f[l + 1]);
        if (di && l == 1 && b == 0) {
          // layer2's first block reads the clip's layer-1 output: frames < D and >= t - D from the clip's edge buffer, the rest
          // in place from the video-level map
          ConvArgs::FrameSel fsel;
          memset(&fsel, 0, sizeof(fsel));
          fsel.on = 1; fsel.a_t = 2 * D; fsel.b_t = t; fsel.b = di->v5; fsel.b_batch_stride = (int64_t)di->fs * F64;
          for (int f = 0; f < t; ++f) {
            const bool edge = f < D || f >= t - D;
            fsel.src[f] = edge ? 0 : 1; fsel.idx[f] = (int16_t)(f < D ? f : edge ? f - (t - 2 * D) : f);
          }
          if ((rc = run_block(h->blocks[l][b], nb, di->e5, ti, hi, wi, out, &ti, &hi, &wi, &fsel, (int64_t)2 * D * F64))) return rc;
        } else {
          if ((rc = run_block(h->blocks[l][b], nb, in, ti, hi, wi, out, &ti, &hi, &wi))) return rc;
        }
        in = out;
      }
    }
    if ((rc = mark(1))) return rc;
    // decoder: lateral projections at native resolution, stem + layer1 share one map
    if (di) {
      // dense-video schedule: the video-level part of this map was projected once per call (run_dense); per clip only its
      // 2 * D edge frames.  Stem output of edge frame j: the clip's own first / last frame, video-level frames otherwise.
      ConvArgs c = make_conv(h->lateral[0], nb, 2 * D, H[0], W[0], di->e1, ws + o_g[0], nullptr, 0, act, 0);
      c.in2 = di->e5; c.macs_per_pos *= 2; c.out_f16 = 1;
      c.in_batch_stride = 2 * F64;
      c.fsel.on = 1; c.fsel.a_t = 2; c.fsel.b_t = t; c.fsel.b = di->v1; c.fsel.b_batch_stride = (int64_t)di->fs * F64;
      for (int j = 0; j < 2 * D; ++j) {
        const int f = j < D ? j : t - 2 * D + j;               // clip frame of edge slot j
        const bool own = f == 0 || f == t - 1;
        c.fsel.src[j] = own ? 0 : 1; c.fsel.idx[j] = (int16_t)(own ? (f == 0 ? 0 : 1) : f);
      }
      if ((rc = run_conv(h, c, stream))) return rc;
    } else {
      ConvArgs c = make_conv(h->lateral[0], nb, T[0], H[0], W[0], ws + o_f[0], ws + o_g[0], nullptr, 0, act, tc_head ? 0 : 1);
      c.in2 = ws + o_f[1];
      c.out_f16 = tc_head ? 1 : 0;               // the tensor-core head reads fp16 lateral maps whatever the trunk's type
      c.macs_per_pos *= 2;
      if ((rc = run_conv(h, c, stream))) return rc;
    }
    for (int i = 2; i < 5; ++i) {
      ConvArgs c = make_conv(h->lateral[i], nb, T[i], H[i], W[i], ws + o_f[i], ws + o_g[i - 1], nullptr, 0, act, tc_head ? 0 : 1);
      c.out_f16 = tc_head ? 1 : 0;
      if ((rc = run_conv(h, c, stream))) return rc;
    }
    if ((rc = mark(2))) return rc;
    HeadArgs ha;
    for (int i = 0; i < 4; ++i) { ha.g[i] = ws + o_g[i]; ha.tl[i] = T[i + 1]; ha.hl[i] = H[i + 1]; ha.wl[i] = W[i + 1]; }
    ha.g0_video = nullptr; ha.g0_lo = ha.g0_hi = ha.g0_step = ha.g0_video_t = 0;
    if (di) { ha.g0_video = di->g0_video; ha.g0_lo = D; ha.g0_hi = t - D; ha.g0_step = di->fs; ha.g0_video_t = di->tv_left; ha.tl[0] = 2 * D; }
    if (tc_head)
      for (int i = 1; i < 4; ++i) {
        ++h->launches;
        if ((rc = launch_temporal_upsample_f16(ws + o_g[i], ws + o_gt[i], nb, T[i + 1], t, H[i + 1], W[i + 1], stream))) return rc;
        ha.g[i] = ws + o_gt[i]; ha.tl[i] = t;
      }
    ha.g_dtype = tc_head ? CLASFV_F16 : CLASFV_F32;
    ha.n = nb; ha.t = t; ha.h = height; ha.w = width;
    ha.b1 = h->b1; ha.w2 = h->w2; ha.b2 = h->b2; ha.wh = h->wh; ha.bh = h->bh;
    ha.a_tab = nullptr; ha.tail_f16 = act == CLASFV_F16 ? 1 : 0;
    if (tc_head && (rc = head_table(h, &ha, stream))) return rc;
    const size_t oes = out_dtype == CLASFV_F32 ? 4 : 2;
    const size_t plane = (size_t)t * height * width * oes;
    ha.seg = seg + (size_t)c0 * (out_kind == CLASFV_OUT_LVPROB ? 1 : 2) * plane; ha.motion = motion + (size_t)c0 * 4 * plane;
    ha.out_dtype = out_dtype; ha.out_kind = out_kind;
    ++h->launches;
    if ((rc = tc_head ? launch_head_umma(ha, h->num_sms, stream) : launch_head(ha, stream))) return rc;
    return mark(3);
  }

  int run() {
    act = h->precision; es = act == CLASFV_F32 ? 4 : 2;
    tc_head = act != CLASFV_F32 && !h->force_simt;      // tensor-core head (reads fp16 lateral maps)
    gs = tc_head ? 2 : 4;
    const int Tn[5] = {t, t, t / 2, t / 4, t / 8};
    const int Hn[5] = {height / 2, height / 2, height / 4, height / 8, height / 16};
    const int Wn[5] = {width / 2, width / 2, width / 4, width / 8, width / 16};
    const int Cn[5] = {64, 64, 128, 256, 512};
    for (int i = 0; i < 5; ++i) { T[i] = Tn[i]; H[i] = Hn[i]; W[i] = Wn[i]; C[i] = Cn[i]; }
    // equally spaced windows of one resident video?
    int64_t frame_step = 0;
    if (offs_host && n >= 2) {
      const int64_t d = offs_host[1] - offs_host[0], hw = (int64_t)height * width;
      bool uniform = d > 0 && d % hw == 0 && d / hw < t;
      for (int i = 2; i < n && uniform; ++i) uniform = offs_host[i] - offs_host[i - 1] == d;
      if (uniform) frame_step = d / hw;
    }
    const bool dense = h->dense_video && tc_head && frame_step >= 1 && frame_step <= 8 && n >= 4 && t >= 16 && t <= CLASFV_MAX_FSEL;
    if (dense) {
      // the video-level maps grow with the run of windows: when they do not fit, the per-clip schedule computes the same
      // (bit-identical) outputs inside the per-batch workspace
      const int rc = run_dense((int)frame_step);
      if (rc != CLASFV_ENOMEM) return rc;
    }
    return run_per_clip(frame_step);
  }

  // ------------------------------------------------------------------ per-clip schedule
  int run_per_clip(int64_t frame_step) {
    const int nbmax = std::min(n, h->sub_batch);
    total = 0;
    carve_batch(nbmax, false);
    int rc;
    if ((rc = ensure_workspace(h, total))) return rc;
    ws = static_cast<char*>(h->ws);
    const int64_t thw = (int64_t)t * height * width;
    for (int c0 = 0; c0 < n; c0 += nbmax) {
      const int nb = std::min(nbmax, n - c0);
      void* offs_dev = nullptr;
      rc = h->ring.upload((size_t)nb * sizeof(int64_t), stream, [&](char* dst) {
        int64_t* o = reinterpret_cast<int64_t*>(dst);
        for (int i = 0; i < nb; ++i) o[i] = offs_host ? offs_host[c0 + i] : (int64_t)(c0 + i) * 3 * thw;
      }, &offs_dev);
      if (rc) return rc;
      if ((rc = mark(-1))) return rc;
      // The 1x7x7 stem convolution is per frame, so equally spaced windows of one video share it: it runs once
      // over the union of their frames and the 3x1x1 convolution that follows reads overlapping windows of that map
      // (clip-edge zero padding comes from the window extent).
      const int64_t fs = nb >= 2 ? frame_step : 0;
      StemArgs sa;
      sa.x = x; sa.clip_offset = static_cast<const int64_t*>(offs_dev); sa.channel_stride = channel_stride;
      sa.n = nb; sa.t = t; sa.h = height; sa.w = width; sa.weight = h->stem_w; sa.bias = h->stem_b; sa.out = ws + o_s0; sa.out_channels = STEM_MID_PAD; sa.out_dtype = act;
      if (fs) { sa.n = 1; sa.t = (int)((nb - 1) * fs + t); }
      ++h->launches;
      if ((rc = launch_stem(sa, stream))) return rc;
      if ((rc = mark(0))) return rc;
      {
        ConvArgs c = make_conv(h->stem_t, nb, T[0], H[0], W[0], ws + o_s0, ws + o_f[0], nullptr, 1, act, 0);
        if (fs) c.in_batch_stride = fs * (int64_t)H[0] * W[0] * STEM_MID_PAD;
        if ((rc = run_conv(h, c, stream))) return rc;
      }
      if ((rc = run_tail(nb, c0, 0))) return rc;
    }
    if (h->profiling) ++h->prof_calls;
    return CLASFV_OK;
  }

  // ------------------------------------------------------------------ dense-video schedule
  int run_dense(int fs) {
    const int nbmax = std::min(n, h->sub_batch);
    const int tv = (n - 1) * fs + t;                       // frames of the video this call touches
    const int64_t FE = (int64_t)H[0] * W[0];
    const int midc = h->blocks[0][0].s1.cout_pad;          // 144
    const int64_t F64 = FE * 64, FM = FE * midc;           // elements per frame
    const int D = DENSE_DEPTH;                             // temporal convolutions through the stem and layer1
    total = 0;
    const size_t o_s0v = region((size_t)tv * FE * STEM_MID_PAD * es);
    size_t o_v[6], o_vs[6];                                // V_d (d = 1..5), VS_d (d = 2..5); V_2 and V_4 share a buffer
    o_v[1] = region((size_t)tv * F64 * es); o_v[3] = region((size_t)tv * F64 * es); o_v[5] = region((size_t)tv * F64 * es);
    o_v[2] = o_v[4] = region((size_t)tv * F64 * es);
    for (int d = 2; d <= D; ++d) o_vs[d] = region((size_t)tv * FM * es);
    size_t o_e[6];
    for (int d = 1; d <= D; ++d) o_e[d] = region((size_t)nbmax * 2 * d * F64 * es);
    const size_t o_g0v = region((size_t)tv * FE * DEC * 2);      // video-level lateral map of level 0 (fp16)
    const size_t o_es = region((size_t)nbmax * 2 * (D - 1) * FM * es);
    carve_batch(nbmax, true);
    int rc;
    if ((rc = ensure_workspace(h, total))) return rc;
    ws = static_cast<char*>(h->ws);

    // ---- video level: stem and layer1 once over the tv frames, as one long clip
    void* offs_dev = nullptr;
    rc = h->ring.upload(sizeof(int64_t), stream, [&](char* dst) { *reinterpret_cast<int64_t*>(dst) = offs_host[0]; }, &offs_dev);
    if (rc) return rc;
    if ((rc = mark(-1))) return rc;
    StemArgs sa;
    sa.x = x; sa.clip_offset = static_cast<const int64_t*>(offs_dev); sa.channel_stride = channel_stride;
    sa.n = 1; sa.t = tv; sa.h = height; sa.w = width; sa.weight = h->stem_w; sa.bias = h->stem_b; sa.out = ws + o_s0v; sa.out_channels = STEM_MID_PAD; sa.out_dtype = act;
    ++h->launches;
    if ((rc = launch_stem(sa, stream))) return rc;
    if ((rc = mark(0))) return rc;
    const PackedConv* spat[6] = {nullptr, nullptr, &h->blocks[0][0].s1, &h->blocks[0][0].s2, &h->blocks[0][1].s1, &h->blocks[0][1].s2};
    const PackedConv* temp[6] = {nullptr, &h->stem_t, &h->blocks[0][0].t1, &h->blocks[0][0].t2, &h->blocks[0][1].t1, &h->blocks[0][1].t2};
    if ((rc = run_conv(h, make_conv(*temp[1], 1, tv, H[0], W[0], ws + o_s0v, ws + o_v[1], nullptr, 1, act, 0), stream))) return rc;
    for (int d = 2; d <= D; ++d) {
      if ((rc = run_conv(h, make_conv(*spat[d], 1, tv, H[0], W[0], ws + o_v[d - 1], ws + o_vs[d], nullptr, 1, act, 0), stream))) return rc;
      const void* res = (d == 3 || d == 5) ? ws + o_v[d - 2] : nullptr;
      if ((rc = run_conv(h, make_conv(*temp[d], 1, tv, H[0], W[0], ws + o_vs[d], ws + o_v[d], res, 1, act, 0), stream))) return rc;
    }
    if ((rc = mark(1))) return rc;
    {
      // level 0 of the decoder's lateral maps over the video: one projection of (stem output, layer1 output) per video
      // frame instead of one per clip frame
      ConvArgs c = make_conv(h->lateral[0], 1, tv, H[0], W[0], ws + o_v[1], ws + o_g0v, nullptr, 0, act, 0);
      c.in2 = ws + o_v[D]; c.macs_per_pos *= 2; c.out_f16 = 1;
      set_stage(2);
      if ((rc = run_conv(h, c, stream))) return rc;
      if ((rc = mark(2))) return rc;
    }

    // ---- per batch of clips: edge frames of layer1, layers 2-4 on virtual clips, decoder
    for (int c0 = 0; c0 < n; c0 += nbmax) {
      const int nb = std::min(nbmax, n - c0);
      if ((rc = mark(-1))) return rc;
      set_stage(1);
      // clip window views of the video-level maps: clip i of the batch starts at frame (c0 + i) * fs
      auto vview = [&](size_t off, int64_t frame_elems) { return ws + off + (size_t)c0 * fs * frame_elems * es; };
      for (int d = 1; d <= D; ++d) {
        const int64_t fin = d == 1 ? FE * STEM_MID_PAD : FM;             // input frame (elements)
        char* e_out = ws + o_e[d];
        const int64_t e_bstride = (int64_t)2 * d * F64;
        const int e_right = d;                                             // first right-edge frame in e_out
        if (d >= 2 &&
            (rc = run_conv(h, make_conv(*spat[d], nb, 2 * (d - 1), H[0], W[0], ws + o_e[d - 1], ws + o_es, nullptr, 1, act, 0), stream))) return rc;
        const char* vsrc = d == 1 ? vview(o_s0v, fin) : vview(o_vs[d], fin);
        const int r = d - 2;                                               // depth of the block input (residual)
        const bool has_res = d == 3 || d == 5;
        // left edge: output frames 0..d-1; virtual input = [edge frames 0..d-2 | video frames d-1, d]
        {
          ConvArgs c = make_conv(*temp[d], nb, d, H[0], W[0], d == 1 ? (const void*)vsrc : (const void*)(ws + o_es), e_out, nullptr, 1, act, 0);
          c.out_batch_stride = e_bstride;
          c.seg.on = 1; c.seg.to = d; c.seg.split = d - 1;
          c.seg.a_t = d == 1 ? t : d - 1; c.seg.a_toff = 0; c.in_batch_stride = d == 1 ? fs * fin : (int64_t)2 * (d - 1) * fin;
          c.seg.b = vsrc; c.seg.b_t = t; c.seg.b_toff = 0; c.seg.b_batch_stride = fs * fin;
          if (has_res) {
            c.residual = ws + o_e[r]; c.res.on = 1; c.res.split = r; c.res.a_toff = 0; c.res.a_batch_stride = (int64_t)2 * r * F64;
            c.res.b = vview(o_v[r], F64); c.res.b_toff = 0; c.res.b_batch_stride = fs * F64;
          }
          if ((rc = run_conv(h, c, stream))) return rc;
        }
        // right edge: output frames t-d..t-1; virtual input = [video frames t-d-1, t-d | edge frames t-d+1..t-1]
        {
          ConvArgs c = make_conv(*temp[d], nb, d, H[0], W[0], vsrc, e_out + (size_t)e_right * F64 * es, nullptr, 1, act, 0);
          c.out_batch_stride = e_bstride;
          c.seg.on = 1; c.seg.to = d; c.seg.split = 1;
          c.seg.a_t = t; c.seg.a_toff = t - d; c.in_batch_stride = fs * fin;
          if (d == 1) { c.seg.b = vsrc; c.seg.b_t = t; c.seg.b_toff = t - d; c.seg.b_batch_stride = fs * fin; }
          else { c.seg.b = ws + o_es + (size_t)(d - 1) * fin * es; c.seg.b_t = d - 1; c.seg.b_toff = -1; c.seg.b_batch_stride = (int64_t)2 * (d - 1) * fin; }
          if (has_res) {
            c.residual = vview(o_v[r], F64); c.res.on = 1; c.res.split = 2; c.res.a_toff = t - d; c.res.a_batch_stride = fs * F64;
            c.res.b = ws + o_e[r] + (size_t)r * F64 * es; c.res.b_toff = -2; c.res.b_batch_stride = (int64_t)2 * r * F64;
          }
          if ((rc = run_conv(h, c, stream))) return rc;
        }
      }
      // No per-clip copy of the stem / layer1 outputs is assembled: layer2 and the lateral projection read virtual clips
      // (ConvArgs::FrameSel), the head reads level 0 from the video-level map between the edges
      DenseIn di;
      di.e1 = ws + o_e[1]; di.e5 = ws + o_e[D];
      di.v1 = vview(o_v[1], F64); di.v5 = vview(o_v[D], F64);
      di.g0_video = ws + o_g0v + (size_t)c0 * fs * FE * DEC * 2;
      di.fs = fs; di.tv_left = tv - c0 * fs;
      if ((rc = run_tail(nb, c0, 1, &di))) return rc;
    }
    if (h->profiling) ++h->prof_calls;
    return CLASFV_OK;
  }
};

}  // namespace

// =================================================================================== C ABI
extern "C" {

int clasfv_abi_version(void) { return CLASFV_ABI_VERSION; }
const char* clasfv_last_error(void) { return g_error; }

int clasfv_create(int device, clasfv_handle** out) {
  if (!out) { set_error("clasfv_create: out is NULL"); return CLASFV_EINVAL; }
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    set_error("clasfv_create: no CUDA device available - this library has no CPU path");
    return CLASFV_EUNSUPPORTED;
  }
  CLASFV_REQUIRE(device >= 0 && device < count, "clasfv_create: device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  CLASFV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("clasfv_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return CLASFV_EUNSUPPORTED;
  }
  DeviceGuard guard(device);
  clasfv_handle* h = new clasfv_handle();
  h->device = device; h->num_sms = prop.multiProcessorCount;
  const char* env = getenv("CLASFV_FORCE_SIMT");
  h->force_simt = env && env[0] == '1';
  int rc = h->ring.init();
  if (rc) { delete h; return rc; }
  *out = h;
  return CLASFV_OK;
}

void clasfv_destroy(clasfv_handle* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  free_packed(h);
  if (h->ws) cudaFree(h->ws);
  if (h->minmax) cudaFree(h->minmax);
  for (auto& kv : h->head_tabs) cudaFree(kv.second);
  for (cudaEvent_t ev : h->prof_events) cudaEventDestroy(ev);
  h->ring.destroy();
  delete h;
}

int clasfv_set_tensor(clasfv_handle* h, const char* key, const float* data_host, const int64_t* shape, int ndim) {
  CLASFV_REQUIRE(h && key && data_host && (shape || ndim == 0) && ndim >= 0 && ndim <= 8, "clasfv_set_tensor: bad argument");
  std::string k(key);
  if (k.rfind("module.", 0) == 0) k = k.substr(7);
  HostTensor t;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) { CLASFV_REQUIRE(shape[i] >= 0, "clasfv_set_tensor: negative extent"); t.shape.push_back(shape[i]); n *= shape[i]; }
  t.data.assign(data_host, data_host + n);
  h->tensors[k] = std::move(t);
  h->finalized = false;
  return CLASFV_OK;
}

int clasfv_finalize(clasfv_handle* h, int precision) {
  CLASFV_REQUIRE(h, "clasfv_finalize: handle is NULL");
  CLASFV_REQUIRE(precision == CLASFV_F32 || precision == CLASFV_BF16 || precision == CLASFV_F16, "clasfv_finalize: unknown precision %d", precision);
  DeviceGuard guard(h->device);
  CLASFV_CUDA(cudaDeviceSynchronize());
  free_packed(h);
  h->precision = precision;
  int rc;
  const std::string p = "r2plus1d_model.";
  // ---- stem: 1x7x7 conv + BN folded, fp32 [147][48]; channels padded to 64 in the activation
  {
    const HostTensor* w;
    if ((rc = need(h, p + "stem.0.weight", {STEM_MID, 3, 1, 7, 7}, &w))) return rc;
    std::vector<float> scale, shift;
    if ((rc = bn_affine(h, p + "stem.1", STEM_MID, &scale, &shift))) return rc;
    std::vector<float> pw(147 * 48, 0.f), pb(48, 0.f);
    for (int co = 0; co < STEM_MID; ++co) {
      for (int tap = 0; tap < 147; ++tap) pw[(size_t)tap * 48 + co] = w->data[(size_t)co * 147 + tap] * scale[co];
      pb[co] = shift[co];
    }
    if ((rc = dev_upload(h, pw.data(), pw.size() * 4, reinterpret_cast<void**>(&h->stem_w)))) return rc;
    if ((rc = dev_upload(h, pb.data(), pb.size() * 4, reinterpret_cast<void**>(&h->stem_b)))) return rc;
    if ((rc = pack_conv_bn(h, p + "stem.3", p + "stem.4", STEM_MID, 64, 3, 1, 1, 1, 1, 1, 1, 0, 0, STEM_MID_PAD, 64, &h->stem_t))) return rc;
  }
  // ---- residual layers
  const int planes_of[4] = {64, 128, 256, 512};
  int inplanes = 64;
  for (int l = 0; l < 4; ++l) {
    const int planes = planes_of[l], stride = l == 0 ? 1 : 2;
    for (int b = 0; b < 2; ++b) {
      const int cin = b == 0 ? inplanes : planes, s = b == 0 ? stride : 1;
      const int mid = midplanes(cin, planes), mid_pad = mid == 230 ? 240 : mid == 460 ? 480 : mid == 921 ? 960 : round_up(mid, 16);
      const std::string k = p + "layer" + std::to_string(l + 1) + "." + std::to_string(b);
      Block& blk = h->blocks[l][b];
      if ((rc = pack_conv_bn(h, k + ".conv1.0.0", k + ".conv1.0.1", cin, mid, 1, 3, 3, 1, s, s, 0, 1, 1, cin, mid_pad, &blk.s1))) return rc;
      if ((rc = pack_conv_bn(h, k + ".conv1.0.3", k + ".conv1.1", mid, planes, 3, 1, 1, s, 1, 1, 1, 0, 0, mid_pad, planes, &blk.t1))) return rc;
      if ((rc = pack_conv_bn(h, k + ".conv2.0.0", k + ".conv2.0.1", planes, mid, 1, 3, 3, 1, 1, 1, 0, 1, 1, planes, mid_pad, &blk.s2))) return rc;
      if ((rc = pack_conv_bn(h, k + ".conv2.0.3", k + ".conv2.1", mid, planes, 3, 1, 1, 1, 1, 1, 1, 0, 0, mid_pad, planes, &blk.t2))) return rc;
      blk.has_down = (b == 0 && stride != 1);
      if (blk.has_down &&
          (rc = pack_conv_bn(h, k + ".downsample.0", k + ".downsample.1", cin, planes, 1, 1, 1, s, s, s, 0, 0, 0, cin, planes, &blk.down))) return rc;
    }
    inplanes = planes;
  }
  // ---- decoder: comb_1 (+BN1) split into five lateral 1x1x1 projections; comb_2 (+BN2); heads
  {
    const HostTensor *w1, *bb1, *w2, *bb2, *ws, *bs, *wm, *bm;
    if ((rc = need(h, "comb_1_layer.weight", {DEC, 1024, 1, 1, 1}, &w1))) return rc;
    if ((rc = need(h, "comb_1_layer.bias", {DEC}, &bb1))) return rc;
    if ((rc = need(h, "comb_2_layer.weight", {DEC, DEC, 1, 1, 1}, &w2))) return rc;
    if ((rc = need(h, "comb_2_layer.bias", {DEC}, &bb2))) return rc;
    if ((rc = need(h, "segmentation_head.weight", {2, DEC, 1, 1, 1}, &ws))) return rc;
    if ((rc = need(h, "segmentation_head.bias", {2}, &bs))) return rc;
    if ((rc = need(h, "motion_head.weight", {4, DEC, 1, 1, 1}, &wm))) return rc;
    if ((rc = need(h, "motion_head.bias", {4}, &bm))) return rc;
    std::vector<float> s1, t1, s2, t2;
    if ((rc = bn_affine(h, "comb_batch_norm_1", DEC, &s1, &t1))) return rc;
    if ((rc = bn_affine(h, "comb_batch_norm_2", DEC, &s2, &t2))) return rc;
    const int widths[5] = {64, 64, 128, 256, 512};
    int off = 0;
    for (int i = 0; i < 5; ++i) {
      PackedConv& pc = h->lateral[i];
      pc = PackedConv();
      pc.cin = pc.cin_pad = widths[i]; pc.cout = pc.cout_pad = DEC;
      if (i == 0) {
        // the stem and layer1 maps share a resolution: one two-source convolution, weights [2][64][64]
        std::vector<float> two((size_t)2 * DEC * 64);
        for (int src = 0; src < 2; ++src)
          for (int co = 0; co < DEC; ++co)
            for (int ci = 0; ci < 64; ++ci) two[((size_t)src * DEC + co) * 64 + ci] = w1->data[(size_t)co * 1024 + src * 64 + ci] * s1[co];
        if (h->precision == CLASFV_F32) {
          if ((rc = dev_upload(h, two.data(), two.size() * 4, &pc.w))) return rc;
        } else if (h->precision == CLASFV_F16) {
          std::vector<__half> th(two.size());
          for (size_t k = 0; k < two.size(); ++k) th[k] = __float2half_rn(two[k]);
          if ((rc = dev_upload(h, th.data(), th.size() * 2, &pc.w))) return rc;
        } else {
          std::vector<__nv_bfloat16> tb(two.size());
          for (size_t k = 0; k < two.size(); ++k) tb[k] = __float2bfloat16_rn(two[k]);
          if ((rc = dev_upload(h, tb.data(), tb.size() * 2, &pc.w))) return rc;
        }
      } else if (i >= 2) {
        if ((rc = pack_weight(h, w1->data.data(), DEC, widths[i], off, 1024, 1, s1.data(), DEC, widths[i], h->precision, &pc.w))) return rc;
      }
      off += widths[i];
    }
    std::vector<float> b1(DEC), w2p(DEC * DEC), b2(DEC), wh(6 * DEC), bh(6);
    for (int j = 0; j < DEC; ++j) {
      b1[j] = s1[j] * bb1->data[j] + t1[j];
      b2[j] = s2[j] * bb2->data[j] + t2[j];
      for (int k = 0; k < DEC; ++k) w2p[j * DEC + k] = w2->data[j * DEC + k] * s2[j];
    }
    for (int k = 0; k < DEC; ++k) {
      wh[0 * DEC + k] = ws->data[k]; wh[1 * DEC + k] = ws->data[DEC + k];
      for (int q = 0; q < 4; ++q) wh[(2 + q) * DEC + k] = wm->data[q * DEC + k];
    }
    bh[0] = bs->data[0]; bh[1] = bs->data[1];
    for (int q = 0; q < 4; ++q) bh[2 + q] = bm->data[q];
    if ((rc = dev_upload(h, b1.data(), b1.size() * 4, reinterpret_cast<void**>(&h->b1)))) return rc;
    if ((rc = dev_upload(h, w2p.data(), w2p.size() * 4, reinterpret_cast<void**>(&h->w2)))) return rc;
    if ((rc = dev_upload(h, b2.data(), b2.size() * 4, reinterpret_cast<void**>(&h->b2)))) return rc;
    if ((rc = dev_upload(h, wh.data(), wh.size() * 4, reinterpret_cast<void**>(&h->wh)))) return rc;
    if ((rc = dev_upload(h, bh.data(), bh.size() * 4, reinterpret_cast<void**>(&h->bh)))) return rc;
  }
  h->finalized = true;
  return CLASFV_OK;
}

int64_t clasfv_workspace_bytes(const clasfv_handle* h) { return h ? (int64_t)h->ws_bytes : 0; }
int64_t clasfv_launch_count(const clasfv_handle* h) { return h ? h->launches : 0; }

int clasfv_forward(clasfv_handle* h, const float* x_dev, const int64_t* clip_offset_host, int64_t channel_stride,
                   int n, int t, int height, int width, int out_kind, int out_dtype,
                   void* seg_dev, void* motion_dev, void* stream_v) {
  CLASFV_REQUIRE(h, "clasfv_forward: handle is NULL");
  if (!h->finalized) { set_error("clasfv_forward: call clasfv_finalize first"); return CLASFV_ESTATE; }
  CLASFV_REQUIRE(x_dev && seg_dev && motion_dev, "clasfv_forward: null buffer");
  CLASFV_REQUIRE(n >= 1 && t >= 8 && t % 8 == 0 && height >= 16 && height % 16 == 0 && width >= 16 && width % 16 == 0,
                 "clasfv_forward: need N >= 1, T %% 8 == 0, H %% 16 == 0, W %% 16 == 0 (got N=%d T=%d H=%d W=%d)", n, t, height, width);
  CLASFV_REQUIRE(out_kind == CLASFV_OUT_LOGITS || out_kind == CLASFV_OUT_PROB || out_kind == CLASFV_OUT_LVPROB, "clasfv_forward: bad out_kind");
  CLASFV_REQUIRE(out_dtype == CLASFV_F32 || out_dtype == CLASFV_BF16 || out_dtype == CLASFV_F16, "clasfv_forward: bad out_dtype");
  DeviceGuard guard(h->device);
  Forward f;
  f.h = h; f.stream = static_cast<cudaStream_t>(stream_v);
  f.x = x_dev; f.offs_host = clip_offset_host; f.n = n; f.t = t; f.height = height; f.width = width;
  f.out_kind = out_kind; f.out_dtype = out_dtype; f.seg = static_cast<char*>(seg_dev); f.motion = static_cast<char*>(motion_dev);
  const int64_t thw = (int64_t)t * height * width;
  if (!clip_offset_host) {
    CLASFV_REQUIRE(channel_stride == 0 || channel_stride == thw, "clasfv_forward: dense input needs channel_stride == T*H*W");
    channel_stride = thw;
  }
  f.channel_stride = channel_stride;
  return f.run();
}

int clasfv_set_option(clasfv_handle* h, const char* name, int value) {
  CLASFV_REQUIRE(h && name, "clasfv_set_option: null argument");
  const std::string k(name);
  if (k == "sub_batch") { CLASFV_REQUIRE(value >= 1 && value <= 4096, "clasfv_set_option: sub_batch out of range"); h->sub_batch = value; }
  else if (k == "dense_video") { h->dense_video = value != 0; }
  else if (k == "umma_pair") { h->umma_pair = value != 0; }
  else { set_error("clasfv_set_option: unknown option '%s'", name); return CLASFV_EINVAL; }
  return CLASFV_OK;
}

int clasfv_profile_begin(clasfv_handle* h) {
  CLASFV_REQUIRE(h, "clasfv_profile_begin: handle is NULL");
  // The event pool is created here, outside any timed region (no driver object creation between kernel launches);
  // 1024 events cover ~70 internal batches, more are created on demand.
  {
    DeviceGuard guard(h->device);
    while (h->prof_events.size() < 1024) {
      cudaEvent_t ev;
      CLASFV_CUDA(cudaEventCreate(&ev));
      h->prof_events.push_back(ev);
    }
    h->prof_stage.reserve(1024);
  }
  h->profiling = true; h->prof_calls = 0; h->prof_used = 0; h->prof_stage.clear();
  for (int s = 0; s < 4; ++s) h->prof_gflop[s] = 0.0;
  return CLASFV_OK;
}

int clasfv_profile_end(clasfv_handle* h, float* stage_ms_host, int* calls_host) {
  CLASFV_REQUIRE(h && stage_ms_host, "clasfv_profile_end: null argument");
  DeviceGuard guard(h->device);
  h->profiling = false;
  for (int s = 0; s < 4; ++s) stage_ms_host[s] = 0.f;
  for (size_t i = 0; i < h->prof_used; ++i) {
    const int stage = h->prof_stage[i];
    if (stage < 0) continue;                       // first event of a call
    CLASFV_CUDA(cudaEventSynchronize(h->prof_events[i]));
    float ms = 0.f;
    CLASFV_CUDA(cudaEventElapsedTime(&ms, h->prof_events[i - 1], h->prof_events[i]));
    stage_ms_host[stage] += ms;
  }
  if (calls_host) *calls_host = h->prof_calls;
  h->prof_calls = 0; h->prof_used = 0; h->prof_stage.clear();
  return CLASFV_OK;
}

int clasfv_profile_gflop(clasfv_handle* h, double* stage_gflop_host) {
  CLASFV_REQUIRE(h && stage_gflop_host, "clasfv_profile_gflop: null argument");
  for (int s = 0; s < 4; ++s) stage_gflop_host[s] = h->prof_gflop[s];
  return CLASFV_OK;
}

int clasfv_ingest_u8(clasfv_handle* h, const uint8_t* frames_dev, int t, int height0, int width0, int bgr,
                     float* video_dev, int height, int width, void* stream_v) {
  CLASFV_REQUIRE(h && frames_dev && video_dev, "clasfv_ingest_u8: null argument");
  CLASFV_REQUIRE(t >= 1 && height0 >= 1 && width0 >= 1 && height >= 1 && width >= 1, "clasfv_ingest_u8: bad extent");
  DeviceGuard guard(h->device);
  if (!h->minmax) CLASFV_CUDA(cudaMalloc(&h->minmax, 6 * sizeof(uint32_t)));
  h->launches += 2;
  return launch_ingest_u8(frames_dev, t, height0, width0, bgr ? 1 : 0, video_dev, height, width, h->minmax, static_cast<cudaStream_t>(stream_v));
}

int clasfv_warp(const float* src_dev, const float* flow_dev, float* out_dev, int n, int c, int height, int width, void* stream) {
  CLASFV_REQUIRE(src_dev && flow_dev && out_dev && n >= 1 && c >= 1 && height >= 1 && width >= 1, "clasfv_warp: bad argument");
  return launch_warp(src_dev, flow_dev, out_dev, n, c, height, width, 0, static_cast<cudaStream_t>(stream));
}

int clasfv_warp_mode(const float* src_dev, const float* flow_dev, float* out_dev, int n, int c, int height, int width, int mode, void* stream) {
  CLASFV_REQUIRE(src_dev && flow_dev && out_dev && n >= 1 && c >= 1 && height >= 1 && width >= 1, "clasfv_warp_mode: bad argument");
  CLASFV_REQUIRE(mode == CLASFV_WARP_BILINEAR || mode == CLASFV_WARP_NEAREST, "clasfv_warp_mode: mode must be CLASFV_WARP_BILINEAR or CLASFV_WARP_NEAREST");
  return launch_warp(src_dev, flow_dev, out_dev, n, c, height, width, mode == CLASFV_WARP_NEAREST ? 1 : 0, static_cast<cudaStream_t>(stream));
}

int clasfv_motion_field(const float* flow_dev, float* grid_dev, int n, int height, int width, void* stream) {
  CLASFV_REQUIRE(flow_dev && grid_dev && n >= 1 && height >= 1 && width >= 1, "clasfv_motion_field: bad argument");
  return launch_motion_field(flow_dev, grid_dev, n, height, width, static_cast<cudaStream_t>(stream));
}

int clasfv_warp_fuse(clasfv_handle* h, const void* prob_dev, int prob_planes, const void* motion_dev, int dtype,
                     const int32_t* clip_start_host, int n_clips, int clip_len, int t_out, int height, int width,
                     int edge_hops, int accumulate, float* acc_dev, int32_t* cnt_dev, uint8_t* mask_dev,
                     int32_t* area_dev, void* stream_v) {
  CLASFV_REQUIRE(h && prob_dev && motion_dev && clip_start_host && acc_dev, "clasfv_warp_fuse: null argument");
  CLASFV_REQUIRE(dtype == CLASFV_F32 || dtype == CLASFV_BF16 || dtype == CLASFV_F16, "clasfv_warp_fuse: bad dtype");
  CLASFV_REQUIRE(n_clips >= 1 && clip_len >= 1 && t_out >= 1 && height >= 1 && width >= 1, "clasfv_warp_fuse: bad extent");
  CLASFV_REQUIRE(prob_planes == 1 || prob_planes == 2, "clasfv_warp_fuse: prob_planes must be 1 (LV) or 2 (background, LV)");
  for (int c = 1; c < n_clips; ++c) CLASFV_REQUIRE(clip_start_host[c] >= clip_start_host[c - 1], "clasfv_warp_fuse: clip starts must ascend");
  DeviceGuard guard(h->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  void* tab = nullptr;
  const size_t bytes = sizeof(int32_t) * ((size_t)n_clips + 2 * (size_t)t_out);
  int rc = h->ring.upload(bytes, stream, [&](char* dst) {
    int32_t* starts = reinterpret_cast<int32_t*>(dst);
    int32_t* lo = starts + n_clips;
    int32_t* hi = lo + t_out;
    memcpy(starts, clip_start_host, sizeof(int32_t) * n_clips);
    // candidate clips of frame g: start in [g - clip_len, g + 1] (direct, forward hop, backward hop, edge hops)
    int a = 0, b = 0;
    for (int g = 0; g < t_out; ++g) {
      while (a < n_clips && clip_start_host[a] < g - clip_len) ++a;
      while (b < n_clips && clip_start_host[b] <= g + 1) ++b;
      lo[g] = a; hi[g] = b;
    }
  }, &tab);
  if (rc) return rc;
  WarpFuseArgs a;
  a.prob = prob_dev; a.motion = motion_dev; a.dtype = dtype; a.prob_planes = prob_planes;
  a.clip_start = static_cast<const int32_t*>(tab); a.frame_lo = a.clip_start + n_clips; a.frame_hi = a.frame_lo + t_out;
  a.n_clips = n_clips; a.clip_len = clip_len; a.t_out = t_out; a.h = height; a.w = width; a.edge_hops = edge_hops; a.accumulate = accumulate;
  a.acc = acc_dev; a.cnt = cnt_dev; a.mask = mask_dev; a.area = area_dev;
  ++h->launches;
  return launch_warp_fuse(a, stream);
}

static int upload_shift_table(clasfv_handle* h, int n_shifts, const int32_t* start, const int32_t* len, const int32_t* nclips,
                              const int32_t* base, int total_clips, cudaStream_t stream, ShiftTable* tab, const int32_t** clip_shift) {
  void* dev = nullptr;
  const size_t bytes = sizeof(int32_t) * (4 * (size_t)n_shifts + (size_t)total_clips);
  int rc = h->ring.upload(bytes, stream, [&](char* dst) {
    int32_t* p = reinterpret_cast<int32_t*>(dst);
    for (int k = 0; k < n_shifts; ++k) {
      p[k] = start ? start[k] : 0; p[n_shifts + k] = len[k]; p[2 * n_shifts + k] = nclips[k]; p[3 * n_shifts + k] = base[k];
      for (int q = 0; q < nclips[k] && base[k] + q < total_clips; ++q) p[4 * n_shifts + base[k] + q] = k;
    }
  }, &dev);
  if (rc) return rc;
  const int32_t* p = static_cast<const int32_t*>(dev);
  tab->start = p; tab->len = p + n_shifts; tab->nclips = p + 2 * n_shifts; tab->clip_base = p + 3 * n_shifts;
  if (clip_shift) *clip_shift = p + 4 * n_shifts;
  return CLASFV_OK;
}

int clasfv_build_shift_clips(clasfv_handle* h, const float* video_dev, int t, int height, int width, int clip_len,
                             int n_shifts, const int32_t* shift_start_host, const int32_t* shift_len_host,
                             const int32_t* shift_nclips_host, const int32_t* shift_clip_base_host,
                             float* clips_dev, void* stream_v) {
  CLASFV_REQUIRE(h && video_dev && clips_dev && shift_start_host && shift_len_host && shift_nclips_host && shift_clip_base_host,
                 "clasfv_build_shift_clips: null argument");
  CLASFV_REQUIRE(n_shifts >= 1 && clip_len >= 1 && t >= 1 && ((int64_t)height * width) % 4 == 0, "clasfv_build_shift_clips: bad extent (H*W must be a multiple of 4)");
  int total = 0;
  for (int k = 0; k < n_shifts; ++k) {
    CLASFV_REQUIRE(shift_start_host[k] >= 0 && shift_len_host[k] >= 1 && shift_start_host[k] + shift_len_host[k] <= t && shift_nclips_host[k] >= 0,
                   "clasfv_build_shift_clips: shift %d out of range", k);
    CLASFV_REQUIRE(shift_clip_base_host[k] == total, "clasfv_build_shift_clips: clip bases must be the running sum of clip counts");
    // without a resample the clips are plain slices and must fit (the reference truncates when it rounds down)
    CLASFV_REQUIRE(shift_nclips_host[k] * clip_len == shift_len_host[k] || shift_len_host[k] % clip_len != 0,
                   "clasfv_build_shift_clips: inconsistent clip count for shift %d", k);
    total += shift_nclips_host[k];
  }
  if (total == 0) return CLASFV_OK;
  DeviceGuard guard(h->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  ShiftTable tab; const int32_t* clip_shift = nullptr;
  int rc = upload_shift_table(h, n_shifts, shift_start_host, shift_len_host, shift_nclips_host, shift_clip_base_host, total, stream, &tab, &clip_shift);
  if (rc) return rc;
  ++h->launches;
  return launch_build_shift_clips(video_dev, t, height, width, clip_len, n_shifts, total, clip_shift, tab, clips_dev, stream);
}

int clasfv_fuse_shift_votes(clasfv_handle* h, const void* prob_dev, int dtype, int t, int height, int width,
                            int clip_len, int step, int n_shifts, const int32_t* shift_len_host,
                            const int32_t* shift_nclips_host, const int32_t* shift_clip_base_host,
                            uint8_t* mask_dev, int32_t* area_dev, void* stream_v) {
  CLASFV_REQUIRE(h && prob_dev && mask_dev && shift_len_host && shift_nclips_host && shift_clip_base_host, "clasfv_fuse_shift_votes: null argument");
  CLASFV_REQUIRE(dtype == CLASFV_F32 || dtype == CLASFV_BF16 || dtype == CLASFV_F16, "clasfv_fuse_shift_votes: bad dtype");
  CLASFV_REQUIRE(n_shifts >= 1 && step >= 1 && clip_len >= 1 && t >= 1, "clasfv_fuse_shift_votes: bad extent");
  CLASFV_REQUIRE(shift_nclips_host[0] >= 1, "clasfv_fuse_shift_votes: shift 0 has no clip (the reference raises IndexError)");
  DeviceGuard guard(h->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  int total = 0;
  for (int k = 0; k < n_shifts; ++k) total += shift_nclips_host[k];
  ShiftTable tab;
  int rc = upload_shift_table(h, n_shifts, nullptr, shift_len_host, shift_nclips_host, shift_clip_base_host, total, stream, &tab, nullptr);
  if (rc) return rc;
  ++h->launches;
  return launch_fuse_shift_votes(prob_dev, dtype, t, height, width, clip_len, step, n_shifts, tab, mask_dev, area_dev, stream);
}

int clasfv_finalize_mask(const float* acc_dev, int t, int height, int width, uint8_t* mask_dev, int32_t* area_dev, void* stream) {
  CLASFV_REQUIRE(acc_dev && (mask_dev || area_dev) && t >= 1 && height >= 1 && width >= 1, "clasfv_finalize_mask: bad argument");
  return launch_finalize_mask(acc_dev, t, height, width, mask_dev, area_dev, static_cast<cudaStream_t>(stream));
}

int clasfv_temporal_resample(const float* in_dev, float* out_dev, int channels, int l_in, int l_out, int64_t hw, void* stream) {
  CLASFV_REQUIRE(in_dev && out_dev && channels >= 1 && l_in >= 1 && l_out >= 1 && hw >= 1, "clasfv_temporal_resample: bad argument");
  return launch_temporal_resample(in_dev, out_dev, channels, l_in, l_out, hw, static_cast<cudaStream_t>(stream));
}

int clasfv_conv3d(clasfv_handle* h, const void* x_dev, int dtype, int n, int t, int height, int width, int cin,
                  const float* w_host, const float* scale_host, const float* shift_host, int cout,
                  int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                  const void* residual_dev, int relu, int engine, int out_f32, void* out_dev, void* stream_v) {
  CLASFV_REQUIRE(h && x_dev && w_host && out_dev, "clasfv_conv3d: null argument");
  CLASFV_REQUIRE(dtype == CLASFV_F32 || dtype == CLASFV_BF16 || dtype == CLASFV_F16, "clasfv_conv3d: bad dtype");
  CLASFV_REQUIRE(cin % 16 == 0 && cout % 16 == 0, "clasfv_conv3d: channel counts must be multiples of 16");
  CLASFV_REQUIRE((engine == 0 && dtype != CLASFV_F16) || (engine == 1 && dtype != CLASFV_F32),
                 "clasfv_conv3d: the tcgen05 engine needs bf16 or fp16, the CUDA-core engine fp32 or bf16");
  DeviceGuard guard(h->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int taps = kt * kh * kw;
  // pack on the host, upload to a temporary that is freed after the stream drains (test surface: simplicity over speed)
  const size_t nw = (size_t)taps * cout * cin;
  std::vector<float> packed(nw);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int tap = 0; tap < taps; ++tap)
        packed[((size_t)tap * cout + co) * cin + ci] = w_host[((size_t)co * cin + ci) * taps + tap] * (scale_host ? scale_host[co] : 1.f);
  void* w_dev = nullptr; float* b_dev = nullptr;
  if (dtype == CLASFV_F32) {
    CLASFV_CUDA(cudaMalloc(&w_dev, nw * 4));
    CLASFV_CUDA(cudaMemcpy(w_dev, packed.data(), nw * 4, cudaMemcpyHostToDevice));
  } else if (dtype == CLASFV_F16) {
    std::vector<__half> ph(nw);
    for (size_t i = 0; i < nw; ++i) ph[i] = __float2half_rn(packed[i]);
    CLASFV_CUDA(cudaMalloc(&w_dev, nw * 2));
    CLASFV_CUDA(cudaMemcpy(w_dev, ph.data(), nw * 2, cudaMemcpyHostToDevice));
  } else {
    std::vector<__nv_bfloat16> pb(nw);
    for (size_t i = 0; i < nw; ++i) pb[i] = __float2bfloat16_rn(packed[i]);
    CLASFV_CUDA(cudaMalloc(&w_dev, nw * 2));
    CLASFV_CUDA(cudaMemcpy(w_dev, pb.data(), nw * 2, cudaMemcpyHostToDevice));
  }
  if (shift_host) {
    CLASFV_CUDA(cudaMalloc(&b_dev, (size_t)cout * 4));
    CLASFV_CUDA(cudaMemcpy(b_dev, shift_host, (size_t)cout * 4, cudaMemcpyHostToDevice));
  }
  PackedConv pc;
  pc.cin = pc.cin_pad = cin; pc.cout = pc.cout_pad = cout;
  pc.kt = kt; pc.kh = kh; pc.kw = kw; pc.st = st; pc.sh = sh; pc.sw = sw; pc.pt = pt; pc.ph = ph; pc.pw = pw;
  pc.w = w_dev; pc.bias = b_dev;
  ConvArgs a = make_conv(pc, n, t, height, width, x_dev, out_dev, residual_dev, relu, dtype, out_f32);
  a.no_pair = h->umma_pair ? 0 : 1;
  int rc = engine == 1 ? launch_conv_umma(a, h->num_sms, stream) : launch_conv_simt(a, stream);
  cudaError_t e = cudaStreamSynchronize(stream);
  cudaFree(w_dev);
  if (b_dev) cudaFree(b_dev);
  if (rc) return rc;
  if (e != cudaSuccess) { set_error("clasfv_conv3d: kernel failed: %s", cudaGetErrorString(e)); return CLASFV_ECUDA; }
  return CLASFV_OK;
}

int clasfv_decoder_head(clasfv_handle* h, const void* g0_dev, const void* g1_dev, const void* g2_dev, const void* g3_dev,
                        int n, int t, int height, int width, int out_kind, int out_dtype,
                        void* seg_dev, void* motion_dev, void* stream_v) {
  CLASFV_REQUIRE(h && g0_dev && g1_dev && g2_dev && g3_dev && seg_dev && motion_dev, "clasfv_decoder_head: null argument");
  if (!h->finalized || h->precision == CLASFV_F32) { set_error("clasfv_decoder_head: finalize the handle in a tensor-core precision first"); return CLASFV_ESTATE; }
  CLASFV_REQUIRE(n >= 1 && t >= 1 && height >= 16 && height % 16 == 0 && width >= 16 && width % 16 == 0, "clasfv_decoder_head: bad extent");
  CLASFV_REQUIRE(out_kind == CLASFV_OUT_LOGITS || out_kind == CLASFV_OUT_PROB || out_kind == CLASFV_OUT_LVPROB, "clasfv_decoder_head: bad out_kind");
  CLASFV_REQUIRE(out_dtype == CLASFV_F32 || out_dtype == CLASFV_BF16 || out_dtype == CLASFV_F16, "clasfv_decoder_head: bad out_dtype");
  DeviceGuard guard(h->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  HeadArgs ha;
  const void* g[4] = {g0_dev, g1_dev, g2_dev, g3_dev};
  for (int i = 0; i < 4; ++i) { ha.g[i] = g[i]; ha.tl[i] = t; ha.hl[i] = height >> (i + 1); ha.wl[i] = width >> (i + 1); }
  ha.g_dtype = CLASFV_F16; ha.n = n; ha.t = t; ha.h = height; ha.w = width;
  ha.b1 = h->b1; ha.w2 = h->w2; ha.b2 = h->b2; ha.wh = h->wh; ha.bh = h->bh;
  ha.a_tab = nullptr; ha.tail_f16 = h->precision == CLASFV_F16 ? 1 : 0;
  ha.g0_video = nullptr; ha.g0_lo = ha.g0_hi = ha.g0_step = ha.g0_video_t = 0;
  ha.seg = seg_dev; ha.motion = motion_dev; ha.out_dtype = out_dtype; ha.out_kind = out_kind;
  int rc = head_table(h, &ha, stream);
  if (rc) return rc;
  return launch_head_umma(ha, h->num_sms, stream);
}

}  // extern "C"
