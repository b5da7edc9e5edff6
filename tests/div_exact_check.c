/* Exhaustive check of the short division used by k_smooth_chunks (csrc/smooth.cuh, div3_small): for y = 1 .. ymax and c = RN(1/y),
 *   q = RN(x c);  r = RN(x - y q) (one FMA, exact);  q' = RN(q + r c)
 * must equal the IEEE quotient x / y for EVERY binary32 significand of x (several exponents inside the range the kernel admits,
 * 2^-100 < |x| < 2^100; the sequence is sign-symmetric).  Test infrastructure: built and run by tests/test_div_exact.py. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(int argc, char** argv)
{
	const int ymax = argc > 1 ? atoi(argv[1]) : 63;
	const int exps[] = { 127, 28 /* 2^-99 */, 226 /* 2^99 */, 100, 150 };
	long bad = 0;
	for (int y = 1; y <= ymax; y++)
	{
		const float fy = (float)y;
		volatile float cv = 1.0f / fy;
		const float c = cv;
		for (int e = 0; e < 5; e++)
			for (uint32_t m = 0; m < (1u << 23); m++)
			{
				const float x = from_bits(((uint32_t)exps[e] << 23) | m);
				const float q = x * c;
				const float r = fmaf(-fy, q, x);
				const float q1 = fmaf(r, c, q);
				volatile float ref = x / fy;
				if (q1 != ref)
				{
					if (bad < 10) printf("y=%d x=%a got %a want %a\n", y, x, q1, (float)ref);
					bad++;
				}
			}
	}
	printf("bad=%ld\n", bad);
	return bad != 0;
}
