/* Exhaustive check of the 3-operation division used by k_circ_peaks / k_circ_stream:
 *   q0 = RN(m*y), e = fma(-q0, d, m), q = fma(e, y, q0)   with y = RN(1/d), d = r*r
 * must equal the IEEE quotient RN(m/d) for every integer |m| <= 2^24 and every r in [1, 24]. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(void)
{
	long bad = 0;
	for (int r = 1; r <= 24; r++) {
		const float d = (float)(r * r), y = 1.0f / d;
		long badr = 0;
#pragma omp parallel for reduction(+ : badr)
		for (int m = -(1 << 24); m <= (1 << 24); m++) {
			const float fm = (float)m, q0 = fm * y, e = fmaf(-q0, d, fm), q1 = fmaf(e, y, q0), ref = fm / d;
			uint32_t a, b;
			memcpy(&a, &q1, 4);
			memcpy(&b, &ref, 4);
			badr += a != b;
		}
		bad += badr;
	}
	printf("%ld\n", bad);
	return bad != 0;
}
