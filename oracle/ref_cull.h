/* TEST INFRASTRUCTURE - CPU restatement of the reference culling / transform hot path.
 *
 * Plain C restatement of nvpro-pipeline's dp::culling::cpu scalar (Linux) path and of
 * dp::transform::Tree::compute, operation for operation.  It exists so
 * that the CUDA path can be checked at sizes where the reference's shared_ptr-per-object
 * API cannot be instantiated (64 M+ objects).  It is NOT the product and NOT a fallback:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) the compiled, unmodified reference (oracle/_ref/libdpref.so) on random scenes incl.
 *       boundary / NaN / Inf / negative-w cases, and
 *   (2) golden vectors generated from that compiled reference (tests/golden/, generator
 *       tools/make_golden.py), incl. the SURVEY.md 8(c) known answer
 *       (3100 visible, 29668 changed, FNV-1a-64 0x0de08815a3db449f).
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction; see oracle/Makefile).
 */
#ifndef DPORACLE_REF_CULL_H
#define DPORACLE_REF_CULL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* r = v * m, row vector times row-major 4x4.            dp/math/Matmnt.h:1371-1379 */
void dporacle_vec4_mul_mat44(const float v[4], const float m[16], float r[4]);

/* r = a * b.                                              dp/math/Matmnt.h:1381-1415 */
void dporacle_mat44_mul(const float a[16], const float b[16], float r[16]);

/* extent = upper - lower in f32 (Box3f::getSize).         dp/math/Boxnt.h:239-242,
 * as used by ManagerBitSet::objectSetBoundingBox          dp/culling/src/ManagerBitSet.cpp:99-103
 * lower/upper/extent are arrays of n float4 (w ignored on input, written 0 for extent). */
void dporacle_box_extent(const float *lower4, const float *upper4, size_t n, float *extent4);

/* OBB of one object: point, ex, ey, ez (16 floats).       dp/culling/cpu/src/ManagerImpl.cpp:114-162 */
void dporacle_obb(const float lower3[3], const float extent3[3], const float m[16], float obb[16]);

/* 8-corner clip-space test of one OBB.                    dp/culling/cpu/src/ManagerImpl.cpp:199-229,263-289
 * returns 1 when visible */
int dporacle_is_visible(const float vp[16], const float obb[16]);

/* Whole cull pass (updateOBBs + visibility loop + BitArray packing).
 *                                                         dp/culling/cpu/src/ManagerImpl.cpp:467-518
 * lower4/extent4: n float4 each (w ignored); transformIndex: n u32; matrices: base pointer
 * + byte stride like GroupBitSet::setMatrices.  words: ceil(n/32) u32, bit i of object i in
 * word i/32 bit i%32, unused tail bits 0 (dp/util/BitArray.h:215-219,298-308). */
void dporacle_cull_bits(const float *lower4, const float *extent4, const uint32_t *transformIndex, size_t n,
                        const void *matrices, size_t strideBytes, const float vp[16], uint32_t *words);

/* same, split over nthreads host threads on 64-object boundaries (bench baseline only) */
void dporacle_cull_bits_mt(const float *lower4, const float *extent4, const uint32_t *transformIndex, size_t n,
                           const void *matrices, size_t strideBytes, const float vp[16], uint32_t *words,
                           int nthreads);

/* ResultBitSet incarnation step: resize the stored result to newN objects, bits of new
 * objects = 1 (visible).                                  dp/culling/src/ResultBitSet.cpp:65-79
 * words must have room for ceil(max(oldN,newN)/32) u32. */
void dporacle_result_resize(uint32_t *resultWords, size_t oldN, size_t newN);

/* ResultBitSet::updateChanged: changed = new ^ stored, list in ascending index, stored = new.
 *                                                         dp/culling/src/ResultBitSet.cpp:100-107
 * returns the number of changed objects written to changedIdx (capacity n). */
size_t dporacle_update_changed(const uint32_t *newWords, uint32_t *resultWords, size_t n, uint32_t *changedIdx);

/* ResultBitSet::onNotify: an object moved oldIndex -> newIndex inside the group.
 *                                                         dp/culling/src/ResultBitSet.cpp:110-128 */
void dporacle_result_move_bit(uint32_t *resultWords, size_t resultSize, size_t oldIndex, size_t newIndex);

/* ManagerBitSet::calculateBoundingBox, scalar branch.     dp/culling/src/ManagerBitSet.cpp:268-306
 * out6 = lower.xyz, upper.xyz */
void dporacle_bounding_box(const float *lower4, const float *extent4, const uint32_t *transformIndex, size_t n,
                           const void *matrices, size_t strideBytes, float out6[6]);

/* dp::transform::Tree::compute.                           dp/transform/src/Tree.cpp:133-166
 * entries: {parent, transform} u32 pairs of all levels back to back, level l occupying
 * entries [levelOffsets[l], levelOffsets[l+1]).  dirtyLocal / dirtyWorld: u32 bit words over
 * node indices.  On return dirtyWorld holds the set the reference publishes through
 * EventWorldMatricesChanged and dirtyLocal is cleared (the reference clears both after
 * notifying; the caller clears dirtyWorld once it has consumed it). */
void dporacle_tree_compute(const float *local, float *world, const uint32_t *entries,
                           const uint32_t *levelOffsets, int numLevels,
                           uint32_t *dirtyLocal, uint32_t *dirtyWorld, size_t numNodes);

/* FNV-1a-64 over per-object visibility bytes in index order (SURVEY.md 8c known answer) */
uint64_t dporacle_visibility_fnv1a(const uint32_t *words, size_t n);

#ifdef __cplusplus
}
#endif
#endif
