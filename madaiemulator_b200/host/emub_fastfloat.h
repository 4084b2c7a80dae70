/*
 * emub_fastfloat.h -- exact decimal -> binary64 conversion for the streaming text protocol, several times faster than
 * strtod on 17-digit input.  Same result as strtod / fscanf("%lf") (round to nearest even) for every token it accepts;
 * anything else (more than 19 significant digits, hex floats, inf / nan, trailing junk, results in the subnormal or
 * overflow range) is declined and the caller uses strtod.
 */
#ifndef EMUB_FASTFLOAT_H
#define EMUB_FASTFLOAT_H
#ifdef __cplusplus
extern "C" {
#endif
/* converts the whole token [p, end); returns 1 and stores the value, or 0 = declined */
int emub_fast_strtod(const char *p, const char *end, double *out);
/* the bytes of printf("%.17f\n", x) (the output format of interactive_mode, interactive_emulator.c:431-437) into buf
 * (at least 48 bytes); returns their number, or 0 = declined (|x| >= 2^63, inf, nan: use snprintf).  Exact: the
 * decimal expansion is rounded to 17 places from the full binary value, ties to even, like glibc. */
int emub_fast_format17(double x, char *buf);
#ifdef __cplusplus
}
#endif
#endif
