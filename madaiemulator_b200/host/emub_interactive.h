/*
 * emub_interactive.h -- interactive_mode (src/interactive_emulator.c:369-450) as a streaming, batched service.
 *
 * The reference reads ONE point with fscanf, emulates it on every PCA component, prints 2*nt lines with "%.17f\n" and
 * fflush()es, per point.  Here stdin is read in large blocks, a block of points goes through emub_predict_multi
 * (all components + back-projection on the GPU), and the answers are written in the identical text format and
 * order, flushed once per block.  The header lines (unless --quiet) are the reference's (:398-414).
 */
#ifndef EMUB_INTERACTIVE_H
#define EMUB_INTERACTIVE_H
#include <stdio.h>
#include "../../include/emu_b200.h"
#include "emub_snapshot.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct emub_multi_emulator emub_multi_emulator;

/* alloc_multi_emulator (multivar_support.c:30-60) from a loaded snapshot: one model carrying the nr component
 * training vectors, one cached factor per component */
int emub_multi_emulator_from_snapshot(emub_ctx *ctx, const emub_snapshot *s, emub_multi_emulator **out);
/* the same on several GPUs of one box: one context + replica per device; emub_multi_emulator_predict (and therefore
 * emub_interactive_stream) then splits every block of query points into contiguous shares, one host thread per
 * device, no exchange between devices (query points are independent).  ndev <= 64. */
int emub_multi_emulator_from_snapshot_devices(const int *devices, int ndev, const emub_snapshot *s, emub_multi_emulator **out);
void emub_multi_emulator_destroy(emub_multi_emulator *me);
int emub_multi_emulator_nt(const emub_multi_emulator *me);
int emub_multi_emulator_nr(const emub_multi_emulator *me);
int emub_multi_emulator_nparams(const emub_multi_emulator *me);
/* emulate_point_multi / _pca for a block: mean, var are m x nt (pca_output: first nr of each row are meaningful) */
int emub_multi_emulator_predict(emub_multi_emulator *me, const double *pts, int m, int pca_output, double *mean, double *var);

/* the same for at most 8 points on the latency path (emub_predict_multi_few): what a point-by-point caller -- the
 * reference's interactive loop, an MCMC driver behind EmuPlusPlus -- should use */
int emub_multi_emulator_predict_few(emub_multi_emulator *me, const double *pts, int m, int pca_output, double *mean, double *var);

/* Text -> doubles, the input side of the stream: buf[0, len) must end on a separator (blank, newline, tab, CR, comma)
 * or be the end of the input, and buf[len] must be writable.  Converts up to max values with strtod (the same
 * conversion as the reference's fscanf("%lf")), tokens counted and converted on `threads` host threads when the text
 * is long.  Returns the number of values; *consumed = bytes used up (the rest belongs to the next block); *bad = 1
 * when a token that is not a number ended the conversion. */
size_t emub_parse_doubles(char *buf, size_t len, double *out, size_t max, int threads, size_t *consumed, int *bad);

/* the interactive_mode loop: reader / device / writer stages on their own host threads over a ring of blocks, text
 * conversion and "%.17f" formatting spread over EMUB_IO_THREADS (default: online cores - 2) worker threads.
 * block_points <= 0 selects 65536 per device (fewer, down to 16384, for models with hundreds of observables).  Returns EMUB_OK at end of input; *npoints (optional)
 * = points answered.  binary != 0 selects the BINARY_INTERACTIVE_MODE framing (raw doubles in and out, :119). */
int emub_interactive_stream(emub_multi_emulator *me, FILE *in, FILE *out, int quiet, int pca_output, int binary,
                            int block_points, long long *npoints);

#ifdef __cplusplus
}
#endif
#endif
