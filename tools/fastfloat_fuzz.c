// fastfloat_fuzz.c -- emub_fast_strtod against strtod on random tokens of a dozen shapes, bit for bit.
// gcc -O2 -std=gnu99 -Imadaiemulator_b200/host -o /tmp/ff_fuzz tools/fastfloat_fuzz.c madaiemulator_b200/host/emub_fastfloat.c -lm && /tmp/ff_fuzz 20000000
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include "emub_fastfloat.h"
static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) { s += 0x9E3779B97F4A7C15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
int main(int argc, char **argv)
{
	long n = argc > 1 ? atol(argv[1]) : 2000000;
	long declined = 0, bad = 0, total = 0;
	char buf[128];
	for (long i = 0; i < n; i++) {
		int kind = (int)(rnd() % 12);
		if (kind == 0) { double v; uint64_t b = rnd(); memcpy(&v, &b, 8); if (!isfinite(v)) continue; snprintf(buf, sizeof buf, "%.17g", v); }
		else if (kind == 1) { double v; uint64_t b = rnd(); memcpy(&v, &b, 8); if (!isfinite(v)) continue; snprintf(buf, sizeof buf, "%.16g", v); }
		else if (kind == 2) { double v = -2.5 + 5.0 * (double)(rnd() >> 11) / 9007199254740992.0; snprintf(buf, sizeof buf, "%.17g", v); }
		else if (kind == 3) { double v = -2.5 + 5.0 * (double)(rnd() >> 11) / 9007199254740992.0; snprintf(buf, sizeof buf, "%.17f", v); }
		else if (kind == 4) { double v = (double)(rnd() >> 11) / 9007199254740992.0; snprintf(buf, sizeof buf, "%.*e", (int)(rnd() % 19), v * pow(10.0, (double)((int)(rnd() % 600) - 300))); }
		else if (kind == 5) { snprintf(buf, sizeof buf, "%llu", (unsigned long long)(rnd() >> (rnd() % 64))); }
		else if (kind == 6) { snprintf(buf, sizeof buf, "%llue%d", (unsigned long long)(rnd() >> (rnd() % 64)), (int)(rnd() % 700) - 350); }
		else if (kind == 7) { snprintf(buf, sizeof buf, "%s0.%0*d%llu", (rnd() & 1) ? "-" : "", (int)(rnd() % 30), 0, (unsigned long long)(rnd() >> (rnd() % 64))); }
		else if (kind == 8) { unsigned long long m = (1ull << 53) + (rnd() % 4096) * 2 + 1; snprintf(buf, sizeof buf, "%llu%s", m, (rnd() & 1) ? "" : "e0"); } /* odd > 2^53: exact ties */
		else if (kind == 9) { unsigned long long m = ((1ull << 52) + rnd() % (1ull << 52)) * 2 + 1; snprintf(buf, sizeof buf, "%llue%d", m, (int)(rnd() % 40) - 20); }
		else if (kind == 10) { snprintf(buf, sizeof buf, "%llu.%llu", (unsigned long long)(rnd() % 100000), (unsigned long long)(rnd() % 10000000000000ull)); }
		else { double v; uint64_t b = (rnd() & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(1 + rnd() % 3) << 52); memcpy(&v, &b, 8); snprintf(buf, sizeof buf, "%.17g", v); } /* near the subnormal boundary */
		total++;
		double a, r = strtod(buf, NULL);
		if (!emub_fast_strtod(buf, buf + strlen(buf), &a)) { declined++; continue; }
		if (memcmp(&a, &r, 8) != 0) { if (bad < 20) printf("MISMATCH %s: fast %.17g (%a) strtod %.17g (%a)\n", buf, a, a, r, r); bad++; }
	}
	printf("%ld tokens, %ld declined (%.2f%%), %ld mismatches\n", total, declined, 100.0 * declined / total, bad);
	return bad != 0;
}
