/*
 * marin_b200.h -- C ABI of the B200-native Marin engine (libmarin_b200.so).
 *
 * Drop-in boundary for PrMers' register engine: every entry point states the reference interface
 * it replaces.  The shape (opaque handle, int return 1 = ok / 0 = failed, thread-local last-error
 * string, nothing thrown across the boundary) follows the reference's own plugin precedent,
 * third_party/aevum/src/EngineApi.h:28-62 (loader: src/aevum/EngineAevum.cpp:89-128).
 *
 * Semantics (include/marin/engine.h:16-303): an engine owns `reg_count` device-resident registers
 * holding residues modulo 2^p - 1.  Calls are asynchronous and in order on the engine's stream; only
 * the get_, equal and sync calls block.  Not thread-safe: one thread per engine.
 * Digits are "encoded": low 32 bits = value, high 32 bits = width of the digit's base.
 */
#ifndef MARIN_B200_H
#define MARIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *mb200_version(void);
const char *mb200_last_error(void);                         /* thread-local; "" if none */

/* ibdwt::transform_size (include/marin/ibdwt.h:17-43); (size_t)-1 if the exponent is too large */
size_t mb200_transform_size_for(uint32_t p);

/* engine::create_gpu(q, reg_count, device, verbose) (src/marin/gpu.cpp:52-150, engine.h:301) */
void *mb200_create(uint32_t p, size_t reg_count, uint32_t device, int verbose);
/* The same with the decomposition of the transform forced (tests and tuning only; every decomposition gives the same
   results): row_length = pairs per row of the contiguous pass (0 = planner's choice; a non-zero value also sends a
   small transform down the multi-pass path), second_level_rows = rows of a second strided pass (0 = only if needed).
   No reference counterpart: the reference picks its kernels per size in a switch (include/marin/engine_gpu.h:238-330). */
void *mb200_create_split(uint32_t p, size_t reg_count, uint32_t device, int verbose, uint32_t row_length, uint32_t second_level_rows);
/* Batched engine (no reference counterpart: the reference creates one engine per ECM curve and runs the curves one
   after the other, src/modes/RunEcmTwistedEdwards.cpp:1736-1755): `batch` independent register sets ("slabs") of
   reg_count registers each, driven in lockstep -- every operation below runs on all slabs, one CTA per slab in the
   same launch.  Single-CTA transform sizes only (n <= 2048: the ECM sizes).  Digits are read / written per slab;
   mb200_get_digits / mb200_set_digits address slab 0; set_u32 writes the same small value to every slab;
   mb200_equal is not available on a batched engine. */
void *mb200_create_batch(uint32_t p, size_t reg_count, uint32_t device, int verbose, uint32_t batch);
uint32_t mb200_batch(void *eng);
int mb200_get_digits_slab(void *eng, uint32_t slab, size_t src, uint64_t *enc, size_t n);
int mb200_set_digits_slab(void *eng, uint32_t slab, size_t dst, const uint64_t *enc, size_t n);
void mb200_destroy(void *eng);                              /* delete eng */
size_t mb200_transform_size(void *eng);                     /* engine::get_size, engine.h:40 */
size_t mb200_register_count(void *eng);
uint32_t mb200_max_multiplier(void *eng);                   /* largest `a` square_mul / mul accept */
int mb200_sync(void *eng);                                  /* engine::sync, engine.h:45 */

int mb200_set_u32(void *eng, size_t dst, uint32_t a);               /* engine::set(Reg, uint32), engine.h:47 */
int mb200_copy(void *eng, size_t dst, size_t src);                  /* engine::copy, engine.h:49 */
int mb200_square_mul(void *eng, size_t r, uint32_t a);              /* engine::square_mul, engine.h:51 */
int mb200_set_multiplicand(void *eng, size_t dst, size_t src);      /* engine::set_multiplicand, engine.h:53 */
/* dst <- multiplicand of (x + y) from the multiplicands a of x and b of y, in one streaming pass (the transform is linear).
   No reference counterpart: src/modes/RunPM1.cpp:3790-3814 forms a V-trace term with copy + add + set_multiplicand; the
   stage-2 driver of this repository uses this instead where the engine allows it.  Needs mb200_max_multiplier() >= 2 (the
   sum has digits of twice the size); the result must only be used as the multiplicand of a mul with a = 1. */
int mb200_add_multiplicand(void *eng, size_t dst, size_t a, size_t b);
int mb200_mul(void *eng, size_t dst, size_t src, uint32_t a);       /* engine::mul, engine.h:60 */
int mb200_sub_u32(void *eng, size_t r, uint32_t a);                 /* engine::sub, engine.h:62 */
int mb200_add(void *eng, size_t dst, size_t src);                   /* engine::add, engine.h:64 */
int mb200_sub_reg(void *eng, size_t dst, size_t src);               /* engine::sub_reg, engine.h:71 */

/* fused helpers with the semantics of the default compositions in engine.h:54-131; each is ONE pass over the registers
   (the reference's overrides: engine_gpu.h:1064-1072,1131; kernels carry_weight_muladd / addsub(_copy) / mul_p1_copy,
   marin.cl:1731-1769,1891-2154,2298-2365).  mul_add fuses for a = 1; outputs of addsub may alias its inputs. */
int mb200_mul_add(void *eng, size_t dst, size_t mul_src, size_t add_src, uint32_t a);       /* engine.h:65-69 */
int mb200_addsub(void *eng, size_t sum_out, size_t diff_out, size_t a, size_t b);           /* engine.h:72-80 */
int mb200_addsub_copy(void *eng, size_t sum, size_t diff, size_t sum_copy, size_t diff_copy, size_t a, size_t b); /* engine.h:124-130 */
int mb200_square_mul_copy(void *eng, size_t r, size_t dst_copy, uint32_t a);                /* engine.h:82-86 */
int mb200_mul_copy(void *eng, size_t dst, size_t src, size_t dst_copy, uint32_t a);         /* engine.h:92-96 */

/* engine::is_equal (engine.h:148-157): residues compared modulo 2^p - 1 (0 and 2^p - 1 are equal); canonical carry and
   comparison run on the device, 16 bytes come back */
int mb200_equal(void *eng, size_t lhs, size_t rhs, int *out);

/* engine::get / engine::set(Reg, uint64*) (engine.h:24-25; engine_gpu.h:1452-1561): canonical digits,
   unweighted and fully carried (on the device; engine_gpu.h:1534-1561 does it on the host); n = transform size entries.
   Every call that returns register contents first reads the device error word and fails if it is set. */
int mb200_get_digits(void *eng, size_t src, uint64_t *enc, size_t n);
int mb200_set_digits(void *eng, size_t dst, const uint64_t *enc, size_t n);

/* engine::get_data / set_data / get_checkpoint / set_checkpoint (engine.h:133-146): raw register
   images, n * 8 bytes each (layout is private to this engine: backend tag 3, never Marin-OpenCL's) */
size_t mb200_register_data_size(void *eng);
int mb200_get_raw(void *eng, size_t src, void *buf, size_t bytes);
int mb200_set_raw(void *eng, size_t dst, const void *buf, size_t bytes);
size_t mb200_checkpoint_size(void *eng);
int mb200_get_checkpoint(void *eng, void *buf, size_t bytes);
int mb200_set_checkpoint(void *eng, const void *buf, size_t bytes);

/* the reference's `-bench` loop (src/core/App.cpp:715-746) on the engine's own stream: `count`
   back-to-back square_mul(r, a), timed with CUDA events and a final synchronise; *ms = elapsed */
int mb200_bench_square_mul(void *eng, size_t r, uint32_t a, uint32_t count, float *ms);
/* same loop without timing (asynchronous) */
int mb200_square_mul_repeat(void *eng, size_t r, uint32_t a, uint32_t count);
/* one kernel of square_mul alone, `count` launches timed with CUDA events (the register content is
   scrambled): which = 0 strided forward pass, 1 contiguous middle pass, 2 strided inverse pass with
   unweight/carry/reweight (carry hand-off between tiles included).  Engines small enough for one CTA only have pass 1. */
int mb200_bench_pass(void *eng, size_t r, int which, uint32_t count, float *ms);
int mb200_is_single_cta(void *eng);

/* Replay of a fixed sequence of engine calls as ONE kernel launch (no reference counterpart: the reference pays one
   clEnqueueNDRangeKernel per op, include/marin/engine_gpu.h:886-1248, which is all an ECM curve at n = 32 costs,
   src/modes/RunEcmTwistedEdwards.cpp:2717-2945).  Between graph_begin and graph_end the asynchronous ops (set_u32,
   copy, square_mul, set_multiplicand, mul, add, sub_reg, sub_u32 and their fused forms) are recorded, not run;
   blocking calls fail.  graph_launch runs the recorded list inside one CTA (k_program): the engine orders the ops by
   their register conflicts only and runs independent ones side by side on the warp groups of the CTA, so the registers
   afterwards are those of the calls made one after the other, in less time (an Edwards doubling of 23 ops: 8 in a chain).
   A register that the list overwrites completely more than once may live in an engine-private register in between;
   its contents DURING a replay are unspecified, its contents afterwards are not.  Single-CTA engines
   (mb200_is_single_cta, n <= 2048: the ECM sizes) only. */
int mb200_graph_begin(void *eng);
int mb200_graph_end(void *eng, int *graph_id);
int mb200_graph_launch(void *eng, int graph_id);
/* kernel launches issued by this engine since creation (bench.py's gpu_launches) */
uint64_t mb200_launch_count(void *eng);

/* free / total bytes of device memory (cudaMemGetInfo): what the P-1 stage-2 planner sizes its baby table with
   (the reference asks OpenCL for CL_DEVICE_GLOBAL_MEM_SIZE, src/modes/RunPM1.cpp:1994-2026) */
int mb200_device_mem_info(uint32_t device, size_t *free_bytes, size_t *total_bytes);

/* pinned host buffers for fast get_/set_ transfers */
void *mb200_host_alloc(size_t bytes);
void mb200_host_free(void *p);

/* The reference's own plugin ABI, exported by the same library (third_party/aevum/src/EngineApi.h:28-62; loader
   src/aevum/EngineAevum.cpp:89-243): with AEVUM_ENGINE_LIB=<path>/libmarin_b200.so the reference's unmodified
   `engine_aevum` adapter (-aevum / engine::gpu_backend::aevum) runs on this engine.  Residues cross as ceil(p / 32)
   packed little-endian 32-bit words; `prepare` is set_multiplicand; fft_spec and tune_dir are ignored. */
const char *aevum_engine_version(void);
const char *aevum_engine_last_error(void);
int aevum_engine_resolve_fft(uint32_t exponent, const char *fft_spec, char *output, size_t output_size);
void *aevum_engine_create(uint32_t exponent, size_t register_count, uint32_t device, int verbose, const char *fft_spec, const char *tune_dir);
void aevum_engine_destroy(void *handle);
size_t aevum_engine_transform_size(void *handle);
size_t aevum_engine_word_count(void *handle);
int aevum_engine_sync(void *handle);
int aevum_engine_set_u32(void *handle, size_t dst, uint32_t value);
int aevum_engine_set_words(void *handle, size_t dst, const uint32_t *words, size_t count);
int aevum_engine_get_words(void *handle, size_t src, uint32_t *words, size_t count);
int aevum_engine_copy(void *handle, size_t dst, size_t src);
int aevum_engine_prepare(void *handle, size_t dst, size_t src);
int aevum_engine_square_mul(void *handle, size_t reg, uint32_t factor);
int aevum_engine_mul(void *handle, size_t dst, size_t src, uint32_t factor);
int aevum_engine_add(void *handle, size_t dst, size_t src);
int aevum_engine_sub_reg(void *handle, size_t dst, size_t src);
int aevum_engine_sub_u32(void *handle, size_t dst, uint32_t value);
int aevum_engine_equal(void *handle, size_t lhs, size_t rhs, int *equal_out);

#ifdef __cplusplus
}
#endif
#endif
