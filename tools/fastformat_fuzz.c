// fastformat_fuzz.c -- emub_fast_format17 against snprintf("%.17f\n") on random doubles of many magnitudes, exact ties included.
// gcc -O2 -std=gnu99 -Imadaiemulator_b200/host -o /tmp/fmt_fuzz tools/fastformat_fuzz.c madaiemulator_b200/host/emub_fastfloat.c -lm && /tmp/fmt_fuzz 20000000
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <time.h>
#include "emub_fastfloat.h"
static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) { s += 0x9E3779B97F4A7C15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
int main(int argc, char **argv)
{
	long n = argc > 1 ? atol(argv[1]) : 2000000;
	long declined = 0, bad = 0;
	char a[64], r[512];
	for (long i = 0; i < n; i++) {
		double x;
		int kind = (int)(rnd() % 8);
		if (kind == 0) { uint64_t b = rnd(); memcpy(&x, &b, 8); }
		else if (kind == 1) x = -2.5 + 5.0 * (double)(rnd() >> 11) / 9007199254740992.0;
		else if (kind == 2) x = ((double)(rnd() % 1000000) + 1.0) * pow(2.0, -(double)(rnd() % 80));          /* dyadic: exact ties */
		else if (kind == 3) x = (double)((rnd() % 100000) * 2 + 1) / 262144.0 * ((rnd() & 1) ? -1.0 : 1.0);     /* odd / 2^18: every one a tie */
		else if (kind == 4) x = (double)(rnd() >> 11) / 9007199254740992.0 * pow(10.0, (double)((int)(rnd() % 40) - 30));
		else if (kind == 5) x = (double)(int64_t)(rnd() >> (rnd() % 64)) * ((rnd() & 1) ? -1.0 : 1.0);
		else if (kind == 6) x = (0.99999999999999999 + (double)(rnd() % 10)) - (double)(rnd() % 3) * 1.1102230246251565e-16;  /* carries into the integer part */
		else { uint64_t b = rnd() & 0x800FFFFFFFFFFFFFull; memcpy(&x, &b, 8); }                                 /* subnormals, zeros */
		int la = emub_fast_format17(x, a);
		int lr = snprintf(r, sizeof r, "%.17f\n", x);
		if (la == 0) { declined++; continue; }
		if (la != lr || memcmp(a, r, (size_t)la) != 0) { if (bad < 20) { a[la] = 0; printf("MISMATCH %a: fast %s snprintf %s", x, a, r); } bad++; }
	}
	printf("%ld values, %ld declined (%.2f%%), %ld mismatches\n", n, declined, 100.0 * declined / n, bad);
	/* speed on the typical case */
	struct timespec t0, t1;
	double acc = 0;
	for (int mode = 0; mode < 2; mode++) {
		clock_gettime(CLOCK_MONOTONIC, &t0);
		for (long i = 0; i < 2000000; i++) {
			double x = -2.5 + 5.0 * (double)((i * 2654435761ull) & 0xFFFFFF) / 16777216.0 + 1e-9 * i;
			if (mode == 0) acc += snprintf(r, sizeof r, "%.17f\n", x); else acc += emub_fast_format17(x, a);
		}
		clock_gettime(CLOCK_MONOTONIC, &t1);
		printf("%s: %.1f ns per value\n", mode ? "fast" : "snprintf", ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / 2e6 * 1e9);
	}
	return bad != 0 || acc < 0;
}
