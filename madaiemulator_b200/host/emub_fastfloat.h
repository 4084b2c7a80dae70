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
#ifdef __cplusplus
}
#endif
#endif
