/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see tfl_reader.h).
 * Minimal TFLite flatbuffer reader.  Field ids follow tensorflow/lite/schema/schema.fbs (v3),
 * the container the reference feeds to tflite::FlatBufferModel (lyra/tflite_model_wrapper.cc:39-44).
 */
#include "tfl_reader.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint16_t rd_u16(const uint8_t* b, size_t o) { uint16_t v; memcpy(&v, b + o, 2); return v; }
static int32_t rd_i32(const uint8_t* b, size_t o) { int32_t v; memcpy(&v, b + o, 4); return v; }
static uint32_t rd_u32(const uint8_t* b, size_t o) { uint32_t v; memcpy(&v, b + o, 4); return v; }

/* absolute offset of field `fid` inside table `tab`, 0 when absent */
static size_t fb_field(const uint8_t* b, size_t tab, int fid) {
  size_t vt = tab - (size_t)(int64_t)rd_i32(b, tab);
  uint16_t vsz = rd_u16(b, vt);
  size_t slot = 4 + 2 * (size_t)fid;
  if (slot >= vsz) return 0;
  uint16_t off = rd_u16(b, vt + slot);
  return off ? tab + off : 0;
}
static size_t fb_indirect(const uint8_t* b, size_t o) { return o + rd_u32(b, o); }
static size_t fb_table(const uint8_t* b, size_t tab, int fid) {
  size_t o = fb_field(b, tab, fid);
  return o ? fb_indirect(b, o) : 0;
}
/* vector field: returns element start, sets *n */
static size_t fb_vec(const uint8_t* b, size_t tab, int fid, uint32_t* n) {
  size_t o = fb_field(b, tab, fid);
  if (!o) { *n = 0; return 0; }
  size_t v = fb_indirect(b, o);
  *n = rd_u32(b, v);
  return v + 4;
}

static int type_size(int t) {
  switch (t) {
    case TFL_F32: case TFL_I32: return 4;
    case TFL_I64: return 8;
    case TFL_U8: case TFL_I8: case TFL_BOOL: return 1;
    default: return 0;
  }
}

tfl_model* tfl_load(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long len = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (len < 16) { fclose(f); return NULL; }
  uint8_t* buf = (uint8_t*)malloc((size_t)len + 16);
  if (fread(buf, 1, (size_t)len, f) != (size_t)len) { fclose(f); free(buf); return NULL; }
  fclose(f);
  if (memcmp(buf + 4, "TFL3", 4) != 0) { free(buf); return NULL; }

  tfl_model* m = (tfl_model*)calloc(1, sizeof(*m));
  m->buf = buf;
  m->len = (size_t)len;
  const uint8_t* b = buf;
  size_t model = fb_indirect(b, 0);

  /* Model.operator_codes (1): OperatorCode{0 deprecated_builtin_code:i8, 3 builtin_code:i32} */
  uint32_t ncodes;
  size_t codes = fb_vec(b, model, 1, &ncodes);
  int* opcode = (int*)calloc(ncodes ? ncodes : 1, sizeof(int));
  for (uint32_t i = 0; i < ncodes; ++i) {
    size_t oc = fb_indirect(b, codes + 4 * i);
    size_t f0 = fb_field(b, oc, 0), f3 = fb_field(b, oc, 3);
    int dep = f0 ? (int8_t)b[f0] : 0;
    int neu = f3 ? rd_i32(b, f3) : 0;
    opcode[i] = dep > neu ? dep : neu;
  }
  /* Model.buffers (4): Buffer{0 data:[u8]} */
  uint32_t nbuf;
  size_t bufs = fb_vec(b, model, 4, &nbuf);

  /* Model.subgraphs (2) */
  uint32_t nsub;
  size_t subs = fb_vec(b, model, 2, &nsub);
  m->nsub = (int)nsub;
  m->sub = (tfl_subgraph*)calloc(nsub, sizeof(tfl_subgraph));
  for (uint32_t s = 0; s < nsub; ++s) {
    tfl_subgraph* sg = &m->sub[s];
    size_t sgt = fb_indirect(b, subs + 4 * s);
    uint32_t n;
    /* SubGraph{0 tensors, 1 inputs, 2 outputs, 3 operators, 4 name} */
    size_t tv = fb_vec(b, sgt, 0, &n);
    sg->ntensors = (int)n;
    sg->tensors = (tfl_tensor*)calloc(n ? n : 1, sizeof(tfl_tensor));
    for (uint32_t i = 0; i < n; ++i) {
      tfl_tensor* t = &sg->tensors[i];
      size_t tt = fb_indirect(b, tv + 4 * i);
      /* Tensor{0 shape:[i32], 1 type:i8, 2 buffer:u32, 3 name, 4 quantization, 5 is_variable} */
      uint32_t nd;
      size_t sh = fb_vec(b, tt, 0, &nd);
      t->ndim = (int)nd;
      t->count = 1;
      for (uint32_t d = 0; d < nd && d < 8; ++d) { t->shape[d] = rd_i32(b, sh + 4 * d); t->count *= (size_t)t->shape[d]; }
      size_t ft = fb_field(b, tt, 1);
      t->type = ft ? (int8_t)b[ft] : 0;
      size_t fbuf = fb_field(b, tt, 2);
      uint32_t bi = fbuf ? rd_u32(b, fbuf) : 0;
      uint32_t nl;
      size_t nm = fb_vec(b, tt, 3, &nl);
      t->name = (const char*)(b + nm);
      t->name_len = (int)nl;
      size_t q = fb_table(b, tt, 4);
      if (q) {
        /* QuantizationParameters{2 scale:[f32], 3 zero_point:[i64], 6 quantized_dimension:i32} */
        uint32_t ns, nz;
        size_t sc = fb_vec(b, q, 2, &ns);
        size_t zp = fb_vec(b, q, 3, &nz);
        t->nscale = (int)ns;
        t->scale = (const float*)(b + sc);
        t->nzp = (int)nz;
        t->zero_point = (const int64_t*)(b + zp);
        size_t qd = fb_field(b, q, 6);
        t->quantized_dimension = qd ? rd_i32(b, qd) : 0;
      }
      if (bi < nbuf) {
        size_t bt = fb_indirect(b, bufs + 4 * bi);
        uint32_t nb;
        size_t d = fb_vec(b, bt, 0, &nb);
        if (nb) { t->data = b + d; t->nbytes = nb; }
      }
      (void)type_size;
    }
    size_t iv = fb_vec(b, sgt, 1, &n); sg->nin = (int)n; sg->inputs = (const int32_t*)(b + iv);
    size_t ov = fb_vec(b, sgt, 2, &n); sg->nout = (int)n; sg->outputs = (const int32_t*)(b + ov);
    size_t opv = fb_vec(b, sgt, 3, &n);
    sg->nops = (int)n;
    sg->ops = (tfl_op*)calloc(n ? n : 1, sizeof(tfl_op));
    for (uint32_t i = 0; i < n; ++i) {
      tfl_op* op = &sg->ops[i];
      size_t ot = fb_indirect(b, opv + 4 * i);
      /* Operator{0 opcode_index:u32, 1 inputs:[i32], 2 outputs:[i32], 4 builtin_options:table} */
      size_t fi = fb_field(b, ot, 0);
      uint32_t oi = fi ? rd_u32(b, fi) : 0;
      op->code = oi < ncodes ? opcode[oi] : -1;
      uint32_t k;
      size_t a = fb_vec(b, ot, 1, &k); op->nin = (int)k; op->in = (const int32_t*)(b + a);
      a = fb_vec(b, ot, 2, &k); op->nout = (int)k; op->out = (const int32_t*)(b + a);
      op->opt = (uint32_t)fb_table(b, ot, 4);
    }
    uint32_t nl;
    size_t nm = fb_vec(b, sgt, 4, &nl);
    sg->name = (const char*)(b + nm);
    sg->name_len = (int)nl;
  }
  /* Model.signature_defs (7): SignatureDef{2 signature_key, 4 subgraph_index} */
  uint32_t nsig;
  size_t sigs = fb_vec(b, model, 7, &nsig);
  m->nsig = (int)nsig;
  m->sig = (tfl_signature*)calloc(nsig ? nsig : 1, sizeof(tfl_signature));
  for (uint32_t i = 0; i < nsig; ++i) {
    size_t st = fb_indirect(b, sigs + 4 * i);
    uint32_t kl;
    size_t k = fb_vec(b, st, 2, &kl);
    m->sig[i].key = (const char*)(b + k);
    m->sig[i].key_len = (int)kl;
    size_t fs = fb_field(b, st, 4);
    m->sig[i].subgraph = fs ? (int)rd_u32(b, fs) : 0;
  }
  free(opcode);
  return m;
}

void tfl_free(tfl_model* m) {
  if (!m) return;
  for (int s = 0; s < m->nsub; ++s) { free(m->sub[s].tensors); free(m->sub[s].ops); }
  free(m->sub);
  free(m->sig);
  free(m->buf);
  free(m);
}

int32_t tfl_opt_i32(const tfl_model* m, const tfl_op* op, int fid, int32_t dflt) {
  if (!op->opt) return dflt;
  size_t o = fb_field(m->buf, op->opt, fid);
  return o ? rd_i32(m->buf, o) : dflt;
}
int8_t tfl_opt_i8(const tfl_model* m, const tfl_op* op, int fid, int8_t dflt) {
  if (!op->opt) return dflt;
  size_t o = fb_field(m->buf, op->opt, fid);
  return o ? (int8_t)m->buf[o] : dflt;
}
float tfl_opt_f32(const tfl_model* m, const tfl_op* op, int fid, float dflt) {
  if (!op->opt) return dflt;
  size_t o = fb_field(m->buf, op->opt, fid);
  if (!o) return dflt;
  float v; memcpy(&v, m->buf + o, 4); return v;
}
int tfl_opt_str(const tfl_model* m, const tfl_op* op, int fid, const char** s) {
  *s = NULL;
  if (!op->opt) return 0;
  uint32_t n;
  size_t v = fb_vec(m->buf, op->opt, fid, &n);
  if (!v) return 0;
  *s = (const char*)(m->buf + v);
  return (int)n;
}

int tfl_signature_subgraph(const tfl_model* m, const char* key) {
  size_t kl = strlen(key);
  for (int i = 0; i < m->nsig; ++i)
    if ((size_t)m->sig[i].key_len == kl && memcmp(m->sig[i].key, key, kl) == 0) return m->sig[i].subgraph;
  return -1;
}
