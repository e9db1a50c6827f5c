/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or
 * executed from the product path (lyra_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Minimal read-only TFLite flatbuffer (schema v3) reader, plain C.
 * The reference loads its three models through tflite::FlatBufferModel::BuildFromFile
 * (reference: lyra/tflite_model_wrapper.cc:39-44); TensorFlow Lite v2.11.0 is an un-vendored
 * dependency (reference: WORKSPACE:168-174), so the container format is restated here from the
 * published schema (tensorflow/lite/schema/schema.fbs, v3): field ids are listed next to each
 * accessor.
 */
#ifndef LYRA_ORACLE_TFL_READER_H_
#define LYRA_ORACLE_TFL_READER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { TFL_F32 = 0, TFL_I32 = 2, TFL_U8 = 3, TFL_I64 = 4, TFL_BOOL = 6, TFL_I8 = 9, TFL_RESOURCE = 13 };

/* builtin operator codes used by the three Lyra graphs */
enum {
  OP_ADD = 0, OP_CONCATENATION = 2, OP_CONV_2D = 3, OP_DEPTHWISE_CONV_2D = 4, OP_DEQUANTIZE = 6,
  OP_MUL = 18, OP_RESHAPE = 22, OP_GATHER = 36, OP_SUB = 41, OP_STRIDED_SLICE = 45, OP_SPLIT = 49,
  OP_CAST = 53, OP_MAXIMUM = 55, OP_ARG_MAX = 56, OP_LESS = 58, OP_TRANSPOSE_CONV = 67,
  OP_NOT_EQUAL = 72, OP_SUM = 74, OP_ARG_MIN = 79, OP_PACK = 83, OP_ONE_HOT = 85, OP_LEAKY_RELU = 98,
  OP_SQUARED_DIFFERENCE = 99, OP_QUANTIZE = 114, OP_CALL_ONCE = 129, OP_VAR_HANDLE = 142,
  OP_READ_VARIABLE = 143, OP_ASSIGN_VARIABLE = 144
};

typedef struct {
  int ndim;
  int shape[8];
  int type;
  size_t count;           /* product of shape (1 for scalars) */
  const char* name;       /* NOT NUL-terminated */
  int name_len;
  int nscale;             /* quantization: per-tensor (1) or per-channel (>1) */
  const float* scale;
  const int64_t* zero_point;
  int nzp;
  int quantized_dimension;
  const uint8_t* data;    /* constant payload or NULL */
  size_t nbytes;
} tfl_tensor;

typedef struct {
  int code;               /* builtin operator code */
  int nin, nout;
  const int32_t* in;
  const int32_t* out;
  uint32_t opt;           /* absolute offset of the builtin_options table, 0 if none */
} tfl_op;

typedef struct {
  int ntensors, nops, nin, nout;
  tfl_tensor* tensors;
  tfl_op* ops;
  const int32_t* inputs;
  const int32_t* outputs;
  const char* name;
  int name_len;
} tfl_subgraph;

typedef struct {
  const char* key;        /* signature_key, NOT NUL-terminated */
  int key_len;
  int subgraph;
} tfl_signature;

typedef struct {
  uint8_t* buf;
  size_t len;
  int nsub;
  tfl_subgraph* sub;
  int nsig;
  tfl_signature* sig;
} tfl_model;

/* returns NULL on any I/O or format error */
tfl_model* tfl_load(const char* path);
void tfl_free(tfl_model* m);

/* builtin_options scalar accessors (field id `fid` of the options table) */
int32_t tfl_opt_i32(const tfl_model* m, const tfl_op* op, int fid, int32_t dflt);
int8_t tfl_opt_i8(const tfl_model* m, const tfl_op* op, int fid, int8_t dflt);
float tfl_opt_f32(const tfl_model* m, const tfl_op* op, int fid, float dflt);
/* string option; returns length, *s points into the file buffer */
int tfl_opt_str(const tfl_model* m, const tfl_op* op, int fid, const char** s);

/* index of the subgraph bound to signature `key`, or -1 */
int tfl_signature_subgraph(const tfl_model* m, const char* key);

#ifdef __cplusplus
}
#endif
#endif
