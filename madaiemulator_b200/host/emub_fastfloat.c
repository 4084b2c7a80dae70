/* emub_fastfloat.c -- see emub_fastfloat.h.  Clinger's exact fast path for small significands and exponents, otherwise
 * the Eisel-Lemire algorithm (D. Lemire, "Number Parsing at a Gigabyte per Second", Software: Practice and Experience
 * 51 (8), 2021) on a 64-bit decimal significand and the 128-bit powers of five of emub_pow5_table.h. */
#include "emub_fastfloat.h"
#include <stdint.h>
#include <string.h>
#include "emub_pow5_table.h"

static const double k_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

static inline void mul64(uint64_t a, uint64_t b, uint64_t *hi, uint64_t *lo)
{
	const unsigned __int128 p = (unsigned __int128)a * b;
	*hi = (uint64_t)(p >> 64);
	*lo = (uint64_t)p;
}

/* w * 10^q, w != 0 a 64-bit significand; returns 0 when the result is not a normal double or the algorithm cannot
 * decide the rounding (the caller falls back to strtod) */
static int eisel_lemire(uint64_t w, int q, int neg, double *out)
{
	if (q < EMUB_POW5_MIN_Q || q > EMUB_POW5_MAX_Q) return 0;
	const int lz = __builtin_clzll(w);
	w <<= lz;
	const unsigned long long *t = emub_pow5_128[q - EMUB_POW5_MIN_Q];
	uint64_t hi, lo;
	mul64(w, t[0], &hi, &lo);
	if ((hi & 0x1FF) == 0x1FF) { /* the 55 bits we need are not settled by the high half of 5^q: bring in the low half */
		uint64_t hi2, lo2;
		mul64(w, t[1], &hi2, &lo2);
		lo += hi2;
		if (hi2 > lo) hi++;
	}
	if (lo == 0xFFFFFFFFFFFFFFFFull && !(q >= -27 && q <= 55)) return 0; /* cannot rule out a carry into the kept bits */
	const int upperbit = (int)(hi >> 63);
	uint64_t mant = hi >> (upperbit + 9);
	int power2 = (int)(((217706 * q) >> 16) + 63) + upperbit - lz + 1023;
	if (power2 <= 0) return 0; /* subnormal: leave it to strtod */
	/* exactly half way between two doubles can only happen for small powers of five: round to even */
	if (lo <= 1 && q >= -4 && q <= 23 && (mant & 3) == 1 && (mant << (upperbit + 9)) == hi) mant &= ~1ull;
	mant += mant & 1;
	mant >>= 1;
	if (mant >= (2ull << 52)) {
		mant = 1ull << 52;
		power2++;
	}
	mant &= ~(1ull << 52);
	if (power2 >= 0x7FF) return 0; /* overflow: strtod sets the range error */
	uint64_t bits = mant | ((uint64_t)power2 << 52) | ((uint64_t)(neg != 0) << 63);
	memcpy(out, &bits, sizeof(bits));
	return 1;
}

/* eight ASCII digits at p -> their value (SWAR); 0 if any of the eight bytes is not a digit */
static inline int eight_digits(const char *p, uint64_t *val)
{
	uint64_t v;
	memcpy(&v, p, 8);
	if (((v & 0xF0F0F0F0F0F0F0F0ull) | (((v + 0x0606060606060606ull) & 0xF0F0F0F0F0F0F0F0ull) >> 4)) != 0x3333333333333333ull) return 0;
	v -= 0x3030303030303030ull;
	v = (v * 10) + (v >> 8);
	v = (((v & 0x000000FF000000FFull) * 0x000F424000000064ull) + (((v >> 16) & 0x000000FF000000FFull) * 0x0000271000000001ull)) >> 32;
	*val = v;
	return 1;
}

int emub_fast_strtod(const char *p, const char *end, double *out)
{
	int neg = 0;
	if (p < end && (*p == '-' || *p == '+')) neg = *p++ == '-';
	uint64_t w = 0;
	int nd = 0;        /* significant digits taken into w */
	int frac = 0;      /* digits after the point, including zeros in front of the first significant one */
	int any = 0;
	while (p < end && *p >= '0' && *p <= '9') {
		if (w != 0 || *p != '0') {
			if (nd == 19) return 0;
			w = w * 10 + (uint64_t)(*p - '0');
			nd++;
		}
		any = 1;
		p++;
	}
	if (p < end && *p == '.') {
		p++;
		while (p < end && *p >= '0' && *p <= '9') {
			uint64_t v8;
			if (w != 0 && nd <= 11 && end - p >= 8 && eight_digits(p, &v8)) { /* every digit is significant from here on */
				w = w * 100000000ull + v8;
				nd += 8;
				frac += 8;
				p += 8;
				continue;
			}
			if (w != 0 || *p != '0') {
				if (nd == 19) return 0;
				w = w * 10 + (uint64_t)(*p - '0');
				nd++;
			}
			frac++;
			any = 1;
			p++;
		}
	}
	if (!any) return 0;
	int ex = 0;
	if (p < end && (*p == 'e' || *p == 'E')) {
		p++;
		int eneg = 0;
		if (p < end && (*p == '-' || *p == '+')) eneg = *p++ == '-';
		if (!(p < end && *p >= '0' && *p <= '9')) return 0;
		while (p < end && *p >= '0' && *p <= '9') {
			if (ex > 100000) return 0;
			ex = ex * 10 + (*p - '0');
			p++;
		}
		if (eneg) ex = -ex;
	}
	if (p != end) return 0; /* trailing characters: strtod decides what they mean */
	if (w == 0) {
		*out = neg ? -0.0 : 0.0;
		return 1;
	}
	const long long q = (long long)ex - frac;
	if (q < -400 || q > 400) return 0;
	/* both operands exact in binary64 and one correctly rounded operation: exact (W. Clinger, 1990) */
	if (w <= (1ull << 53) && q >= -22 && q <= 22) {
		double v = (double)w;
		v = q < 0 ? v / k_pow10[-q] : v * k_pow10[q];
		*out = neg ? -v : v;
		return 1;
	}
	return eisel_lemire(w, (int)q, neg, out);
}

/* ---- "%.17f\n" ------------------------------------------------------------------------------------------------ */
static const char k_digit_pairs[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";

/* exactly ndig decimal digits of v (zero padded) */
static inline void put_digits(char *b, uint64_t v, int ndig)
{
	int i = ndig;
	while (i >= 2) {
		const unsigned r = (unsigned)(v % 100);
		v /= 100;
		i -= 2;
		b[i] = k_digit_pairs[2 * r];
		b[i + 1] = k_digit_pairs[2 * r + 1];
	}
	if (i == 1) b[0] = (char)('0' + (int)(v % 10));
}

int emub_fast_format17(double x, char *buf)
{
	uint64_t bits;
	memcpy(&bits, &x, sizeof(bits));
	const int neg = (int)(bits >> 63);
	const int bexp = (int)((bits >> 52) & 0x7FF);
	uint64_t m = bits & 0x000FFFFFFFFFFFFFull;
	if (bexp == 0x7FF) return 0;       /* inf, nan */
	int e;                             /* x = m * 2^e */
	if (bexp == 0) e = -1074;          /* subnormal (or zero) */
	else { m |= 1ull << 52; e = bexp - 1075; }
	if (e > 10) return 0;              /* |x| >= 2^63 */
	uint64_t ip, q;
	if (e >= 0) { ip = m << e; q = 0; }
	else {
		const int s = -e;              /* 1 .. 1074 */
		uint64_t fm;
		if (s < 64) { ip = m >> s; fm = m & ((1ull << s) - 1); }
		else { ip = 0; fm = m; }
		if (s >= 128) q = 0;           /* fm * 10^17 < 2^110: rounds to 0 at 17 places */
		else {
			const unsigned __int128 P = (unsigned __int128)fm * 100000000000000000ull;
			const unsigned __int128 one = 1;
			const unsigned __int128 quo = P >> s, rem = P & ((one << s) - 1), half = one << (s - 1);
			q = (uint64_t)quo;
			if (rem > half || (rem == half && (q & 1))) q++;
			if (q == 100000000000000000ull) { q = 0; ip++; }
		}
	}
	char *b = buf;
	if (neg) *b++ = '-';
	/* integer part: up to 19 digits, no leading zeros */
	char tmp[20];
	int n = 0;
	do { tmp[n++] = (char)('0' + (int)(ip % 10)); ip /= 10; } while (ip);
	while (n) *b++ = tmp[--n];
	*b++ = '.';
	put_digits(b, q / 1000000000ull, 8);
	put_digits(b + 8, q % 1000000000ull, 9);
	b += 17;
	*b++ = '\n';
	return (int)(b - buf);
}
