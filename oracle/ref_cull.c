/* TEST INFRASTRUCTURE - see ref_cull.h.  CPU restatement of the reference hot path; never
 * linked into, loaded by or substituted for the product (pipeline_b200/lib/libdpcu.so).
 *
 * Every function names the reference lines it restates.  Arithmetic is IEEE binary32,
 * one rounding per operation, evaluated strictly left to right like the reference's
 * expressions; compile with -ffp-contract=off. */
#include "ref_cull.h"

#include <float.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* dp/math/Matmnt.h:1371-1379 : each component is ((v0*m0j + v1*m1j) + v2*m2j) + v3*m3j */
void dporacle_vec4_mul_mat44(const float v[4], const float m[16], float r[4])
{
    for (int j = 0; j < 4; ++j) {
        float acc = v[0] * m[0 * 4 + j];
        acc = acc + v[1] * m[1 * 4 + j];
        acc = acc + v[2] * m[2 * 4 + j];
        acc = acc + v[3] * m[3 * 4 + j];
        r[j] = acc;
    }
}

/* dp/math/Matmnt.h:1381-1415 : row i of the product is (row i of a) * b, same association */
void dporacle_mat44_mul(const float a[16], const float b[16], float r[16])
{
    float t[16];
    for (int i = 0; i < 4; ++i) {
        dporacle_vec4_mul_mat44(a + 4 * i, b, t + 4 * i);
    }
    memcpy(r, t, sizeof t);
}

/* dp/culling/src/ManagerBitSet.cpp:99-103 with Boxnt::getSize (dp/math/Boxnt.h:239-242) */
void dporacle_box_extent(const float *lower4, const float *upper4, size_t n, float *extent4)
{
    for (size_t i = 0; i < n; ++i) {
        extent4[4 * i + 0] = upper4[4 * i + 0] - lower4[4 * i + 0];
        extent4[4 * i + 1] = upper4[4 * i + 1] - lower4[4 * i + 1];
        extent4[4 * i + 2] = upper4[4 * i + 2] - lower4[4 * i + 2];
        extent4[4 * i + 3] = 0.0f;
    }
}

/* dp/culling/cpu/src/ManagerImpl.cpp:127-153 : point = (lower,1) * M ; e{x,y,z} = extent[k] * M[k]
 * (scalar times the full 4-component matrix row, dp/math/Vecnt.h:907-921) */
void dporacle_obb(const float lower3[3], const float extent3[3], const float m[16], float obb[16])
{
    const float l4[4] = { lower3[0], lower3[1], lower3[2], 1.0f };
    dporacle_vec4_mul_mat44(l4, m, obb);
    for (int k = 0; k < 3; ++k) {
        for (int c = 0; c < 4; ++c) {
            obb[4 + 4 * k + c] = m[4 * k + c] * extent3[k];
        }
    }
}

/* dp/culling/cpu/src/ManagerImpl.cpp:199-229 */
static inline unsigned cull_flags(const float p[4])
{
    unsigned cf = 0;
    if (p[0] <= -p[3])      cf |= 0x01u;
    else if (p[3] <= p[0])  cf |= 0x02u;
    if (p[1] <= -p[3])      cf |= 0x04u;
    else if (p[3] <= p[1])  cf |= 0x08u;
    if (p[2] <= -p[3])      cf |= 0x10u;
    else if (p[3] <= p[2])  cf |= 0x20u;
    return cf;
}

static inline void vec4_add(const float a[4], const float b[4], float r[4])
{
    r[0] = a[0] + b[0];
    r[1] = a[1] + b[1];
    r[2] = a[2] + b[2];
    r[3] = a[3] + b[3];
}

/* dp/culling/cpu/src/ManagerImpl.cpp:263-289 */
int dporacle_is_visible(const float vp[16], const float obb[16])
{
    float v[8][4], x[4], y[4], z[4];
    dporacle_vec4_mul_mat44(obb + 0, vp, v[0]);
    dporacle_vec4_mul_mat44(obb + 4, vp, x);
    dporacle_vec4_mul_mat44(obb + 8, vp, y);
    dporacle_vec4_mul_mat44(obb + 12, vp, z);
    vec4_add(v[0], x, v[1]);
    vec4_add(v[0], y, v[2]);
    vec4_add(v[1], y, v[3]);
    vec4_add(v[0], z, v[4]);
    vec4_add(v[1], z, v[5]);
    vec4_add(v[2], z, v[6]);
    vec4_add(v[3], z, v[7]);
    unsigned cfo = 0u, cfa = ~0u;
    for (int i = 0; i < 8; ++i) {
        unsigned cf = cull_flags(v[i]);
        cfo |= cf;
        cfa &= cf;
    }
    return (!cfo || !cfa) ? 1 : 0;
}

static void cull_range(const float *lower4, const float *extent4, const uint32_t *tidx, size_t begin, size_t end,
                       const char *mats, size_t stride, const float vp[16], uint32_t *words)
{
    for (size_t i = begin; i < end; ++i) {
        float obb[16];
        const float *m = (const float *)(mats + (size_t)tidx[i] * stride);
        dporacle_obb(lower4 + 4 * i, extent4 + 4 * i, m, obb);
        if (dporacle_is_visible(vp, obb)) {
            words[i >> 5] |= 1u << (i & 31);
        }
    }
}

/* dp/culling/cpu/src/ManagerImpl.cpp:467-518 (bit layout dp/util/BitArray.h:215-219) */
void dporacle_cull_bits(const float *lower4, const float *extent4, const uint32_t *transformIndex, size_t n,
                        const void *matrices, size_t strideBytes, const float vp[16], uint32_t *words)
{
    memset(words, 0, ((n + 31) / 32) * sizeof(uint32_t));
    cull_range(lower4, extent4, transformIndex, 0, n, (const char *)matrices, strideBytes, vp, words);
}

struct cull_job {
    const float *lower4, *extent4;
    const uint32_t *tidx;
    size_t begin, end;
    const char *mats;
    size_t stride;
    const float *vp;
    uint32_t *words;
};

static void *cull_job_main(void *p)
{
    struct cull_job *j = (struct cull_job *)p;
    cull_range(j->lower4, j->extent4, j->tidx, j->begin, j->end, j->mats, j->stride, j->vp, j->words);
    return NULL;
}

void dporacle_cull_bits_mt(const float *lower4, const float *extent4, const uint32_t *transformIndex, size_t n,
                           const void *matrices, size_t strideBytes, const float vp[16], uint32_t *words,
                           int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    memset(words, 0, ((n + 31) / 32) * sizeof(uint32_t));
    size_t chunks = (n + 63) / 64;                       /* slices start on 64-object boundaries */
    size_t per = (chunks + (size_t)nthreads - 1) / (size_t)nthreads;
    pthread_t tid[256];
    struct cull_job job[256];
    int started = 0;
    for (int t = 0; t < nthreads; ++t) {
        size_t b = (size_t)t * per * 64, e = b + per * 64;
        if (b >= n) break;
        if (e > n) e = n;
        job[t] = (struct cull_job){ lower4, extent4, transformIndex, b, e, (const char *)matrices, strideBytes, vp, words };
        pthread_create(&tid[t], NULL, cull_job_main, &job[t]);
        ++started;
    }
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
}

/* dp/culling/src/ResultBitSet.cpp:65-79 with BitArray::resize (dp/util/src/BitArray.cpp:64-92):
 * surviving bits keep their value, bits of new objects become 1, bits past newN are 0 */
void dporacle_result_resize(uint32_t *w, size_t oldN, size_t newN)
{
    for (size_t i = oldN; i < newN; ++i) w[i >> 5] |= 1u << (i & 31);
    size_t oldWords = (oldN + 31) / 32, newWords = (newN + 31) / 32;
    if (newN % 32) w[newWords - 1] &= ~0u >> (32 - newN % 32);
    for (size_t k = newWords; k < oldWords; ++k) w[k] = 0u;
}

/* dp/culling/src/ResultBitSet.cpp:100-107 ; order = BitArray::traverseBits, ascending
 * (dp/util/BitArray.h:104-136) */
size_t dporacle_update_changed(const uint32_t *newWords, uint32_t *resultWords, size_t n, uint32_t *changedIdx)
{
    size_t count = 0, nw = (n + 31) / 32;
    for (size_t k = 0; k < nw; ++k) {
        uint32_t nv = newWords[k];
        if (k == nw - 1 && (n % 32)) nv &= ~0u >> (32 - n % 32);   /* setBits clears the unused tail */
        uint32_t diff = nv ^ resultWords[k];
        for (unsigned b = 0; diff; ++b, diff >>= 1) {
            if (diff & 1u) changedIdx[count++] = (uint32_t)(k * 32 + b);
        }
        resultWords[k] = nv;
    }
    return count;
}

/* dp/culling/src/ResultBitSet.cpp:110-128 */
void dporacle_result_move_bit(uint32_t *w, size_t resultSize, size_t oldIndex, size_t newIndex)
{
    if (newIndex < resultSize) {
        int value = 1;                                             /* unknown source: assume visible */
        if (oldIndex < resultSize) value = (int)((w[oldIndex >> 5] >> (oldIndex & 31)) & 1u);
        if (value) w[newIndex >> 5] |= 1u << (newIndex & 31);
        else       w[newIndex >> 5] &= ~(1u << (newIndex & 31));
    }
}

/* Boxnt::update, dp/math/Boxnt.h:244-258 */
static inline void box_update(float *lo, float *hi, const float *p, int dims)
{
    for (int i = 0; i < dims; ++i) {
        if (lo[i] > p[i]) lo[i] = p[i];
        if (hi[i] < p[i]) hi[i] = p[i];
    }
}

/* dp/culling/src/ManagerBitSet.cpp:268-306 (scalar branch): Box4f over the 8 world-space
 * corners of every object; result re-wrapped in Box3f(lower.xyz, upper.xyz), whose
 * constructor is init()+update()+update() (dp/math/Boxnt.h:215-221) */
void dporacle_bounding_box(const float *lower4, const float *extent4, const uint32_t *transformIndex, size_t n,
                           const void *matrices, size_t strideBytes, float out6[6])
{
    float lo[4] = { FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX };
    float hi[4] = { -FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (size_t i = 0; i < n; ++i) {
        const float *m = (const float *)((const char *)matrices + (size_t)transformIndex[i] * strideBytes);
        float obb[16], v[8][4];
        dporacle_obb(lower4 + 4 * i, extent4 + 4 * i, m, obb);
        memcpy(v[0], obb, sizeof v[0]);
        vec4_add(v[0], obb + 4, v[1]);
        vec4_add(v[0], obb + 8, v[2]);
        vec4_add(v[1], obb + 8, v[3]);
        vec4_add(v[0], obb + 12, v[4]);
        vec4_add(v[1], obb + 12, v[5]);
        vec4_add(v[2], obb + 12, v[6]);
        vec4_add(v[3], obb + 12, v[7]);
        for (int k = 0; k < 8; ++k) box_update(lo, hi, v[k], 4);
    }
    float blo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, bhi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    box_update(blo, bhi, lo, 3);
    box_update(blo, bhi, hi, 3);
    memcpy(out6, blo, sizeof blo);
    memcpy(out6 + 3, bhi, sizeof bhi);
}

static inline int get_bit(const uint32_t *w, size_t i) { return (int)((w[i >> 5] >> (i & 31)) & 1u); }

/* dp/transform/src/Tree.cpp:133-166 */
void dporacle_tree_compute(const float *local, float *world, const uint32_t *entries,
                           const uint32_t *levelOffsets, int numLevels,
                           uint32_t *dirtyLocal, uint32_t *dirtyWorld, size_t numNodes)
{
    for (int l = 0; l < numLevels; ++l) {
        for (uint32_t e = levelOffsets[l]; e < levelOffsets[l + 1]; ++e) {
            uint32_t parent = entries[2 * e + 0], t = entries[2 * e + 1];
            if (get_bit(dirtyWorld, parent) || get_bit(dirtyLocal, t)) {
                dporacle_mat44_mul(local + 16 * (size_t)t, world + 16 * (size_t)parent, world + 16 * (size_t)t);
                dirtyWorld[t >> 5] |= 1u << (t & 31);
            }
        }
    }
    memset(dirtyLocal, 0, ((numNodes + 31) / 32) * sizeof(uint32_t));
}

uint64_t dporacle_visibility_fnv1a(const uint32_t *words, size_t n)
{
    uint64_t h = 14695981039346656037ull;
    for (size_t i = 0; i < n; ++i) {
        uint64_t v = (words[i >> 5] >> (i & 31)) & 1u;
        h = (h ^ v) * 1099511628211ull;
    }
    return h;
}
