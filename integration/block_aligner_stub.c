/* integration/block_aligner_stub.c -- link-time stand-in for the Rust block-aligner C API (lib/block-aligner/c/block_aligner.h,
 * crate block_aligner_c 0.4.0), used ONLY to build the reference host in an image without cargo/rustc.
 * Every aligner reports a failed alignment (score -1e9), so StripedSmithWaterman.cpp:871-882 takes the reference's own
 * fallback (alignStartPosBacktrace) for word-mode start positions / CIGARs.  Both binaries of the wall-clock
 * comparison (AVX2 baseline and B200 drop-in) are linked against this same stub.  The 32 symbols below are the ones src/
 * references (StripedSmithWaterman.cpp, BlockAligner.cpp). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "block_aligner.h"

static void *blob(void) { return calloc(1, 256); }
BlockHandle block_new_aa_trace_xdrop(uintptr_t a, uintptr_t b, uintptr_t c) { (void) a; (void) b; (void) c; return blob(); }
void block_free_aa_trace_xdrop(BlockHandle b) { free(b); }
BlockHandle block_new_aa_xdrop(uintptr_t a, uintptr_t b, uintptr_t c) { (void) a; (void) b; (void) c; return blob(); }
void block_free_aa_xdrop(BlockHandle b) { free(b); }
struct PaddedBytes *block_new_padded_aa(uintptr_t a, uintptr_t b) { (void) a; (void) b; return (struct PaddedBytes *) blob(); }
void block_free_padded_aa(struct PaddedBytes *p) { free(p); }
struct PosBias *block_new_pos_bias(uintptr_t a, uintptr_t b) { (void) a; (void) b; return (struct PosBias *) blob(); }
void block_free_pos_bias(struct PosBias *p) { free(p); }
struct AAMatrix *block_new_simple_aamatrix(int8_t a, int8_t b) { (void) a; (void) b; return (struct AAMatrix *) blob(); }
void block_free_aamatrix(struct AAMatrix *m) { free(m); }
void block_set_aamatrix_num(struct AAMatrix *m, int8_t a, int8_t b, int8_t s) { (void) m; (void) a; (void) b; (void) s; }
/* the profile branch of alignStartPosBacktraceBlock writes the object's tables itself (StripedSmithWaterman.cpp:964-992) */
struct AAProfile *block_new_aaprofile(uintptr_t len, uintptr_t bs, int8_t ge) { (void) bs; (void) ge; return (struct AAProfile *) calloc(1, (len + 2) * 64 + 4096); }
void block_free_aaprofile(struct AAProfile *p) { free(p); }
size_t block_get_curr_len_aaprofile(const struct AAProfile *p) { (void) p; return 0; }
void block_set_all_gap_open_C_aaprofile(struct AAProfile *p, int8_t g) { (void) p; (void) g; }
void block_set_all_gap_close_C_aaprofile(struct AAProfile *p, int8_t g) { (void) p; (void) g; }
void block_set_all_gap_open_R_aaprofile(struct AAProfile *p, int8_t g) { (void) p; (void) g; }
int8_t *aaprofile_pos_aa(struct AAProfile *p) { return (int8_t *) p; }
int16_t *aaprofile_aa_pos(struct AAProfile *p) { return (int16_t *) p; }
void block_set_bytes_padded_aa(struct PaddedBytes *p, const uint8_t *s, uintptr_t l, uintptr_t m) { (void) p; (void) s; (void) l; (void) m; }
void block_set_bytes_padded_aa_numsequence(struct PaddedBytes *p, const uint8_t *s, uintptr_t l, uintptr_t m) { (void) p; (void) s; (void) l; (void) m; }
void block_set_pos_bias(struct PosBias *p, const int16_t *b, uintptr_t l) { (void) p; (void) b; (void) l; }
void block_align_aa_trace_xdrop_posbias(BlockHandle b, const struct PaddedBytes *q, const struct PosBias *qb, const struct PaddedBytes *r,
                                        const struct PosBias *rb, const struct AAMatrix *m, struct Gaps g, struct SizeRange s, int32_t x) {
    (void) b; (void) q; (void) qb; (void) r; (void) rb; (void) m; (void) g; (void) s; (void) x;
}
void block_align_aa_xdrop_posbias(BlockHandle b, const struct PaddedBytes *q, const struct PosBias *qb, const struct PaddedBytes *r,
                                  const struct PosBias *rb, const struct AAMatrix *m, struct Gaps g, struct SizeRange s, int32_t x) {
    (void) b; (void) q; (void) qb; (void) r; (void) rb; (void) m; (void) g; (void) s; (void) x;
}
void block_align_profile_aa_trace_xdrop(BlockHandle b, const struct PaddedBytes *q, const struct AAProfile *p, struct SizeRange s, int32_t x) {
    (void) b; (void) q; (void) p; (void) s; (void) x;
}
static struct AlignResult failed(void) { struct AlignResult r; memset(&r, 0, sizeof(r)); r.score = -1000000000; return r; }
struct AlignResult block_res_aa_trace_xdrop(BlockHandle b) { (void) b; return failed(); }
struct AlignResult block_res_aa_xdrop(BlockHandle b) { (void) b; return failed(); }
struct Cigar *block_new_cigar(uintptr_t a, uintptr_t b) { (void) a; (void) b; return (struct Cigar *) blob(); }
void block_free_cigar(struct Cigar *c) { free(c); }
void block_cigar_aa_trace_xdrop(BlockHandle b, uintptr_t i, uintptr_t j, struct Cigar *c) { (void) b; (void) i; (void) j; (void) c; }
uintptr_t block_len_cigar(const struct Cigar *c) { (void) c; return 0; }
struct OpLen block_get_cigar(const struct Cigar *c, uintptr_t i) { struct OpLen o; (void) c; (void) i; memset(&o, 0, sizeof(o)); return o; }
