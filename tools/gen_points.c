/* gen_points START COUNT D [binary]: query points for the streaming bench, uniform in [-2.5, 2.5]^D from a counter-based
 * generator (splitmix64 of the value index), "%.17g" text (one point per line) or raw doubles. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
static uint64_t splitmix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}
int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: %s START COUNT D [binary]\n", argv[0]); return 2; }
	const long long start = atoll(argv[1]), count = atoll(argv[2]);
	const int d = atoi(argv[3]), binary = argc > 4;
	static char obuf[1 << 20];
	setvbuf(stdout, obuf, _IOFBF, sizeof(obuf));
	for (long long q = start; q < start + count; q++) {
		for (int k = 0; k < d; k++) {
			const double u = (double)(splitmix64(0x5eed0000ull + (uint64_t)q * (uint64_t)d + (uint64_t)k) >> 11) * (1.0 / 9007199254740992.0);
			const double v = -2.5 + 5.0 * u;
			if (binary) fwrite(&v, sizeof(v), 1, stdout);
			else printf(k + 1 < d ? "%.17g " : "%.17g\n", v);
		}
	}
	return 0;
}
