/* Exhaustive check of rc_device.cuh div32767 (the FCHK-free division used by oct_decode): for every int16 value the Newton
 * sequence equals the correctly rounded a / 32767.0f bit for bit.  Built and run by tests/test_abi.py. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static float div32767(float a)
{
    const float r0 = 3.0518509447574615479e-05f;
    const float r = fmaf(fmaf(r0, -32767.0f, 1.0f), r0, r0);
    const float q = a * r;
    return fmaf(r, fmaf(q, -32767.0f, a), q);
}

int main(void)
{
    int bad = 0;
    for (int i = -32768; i <= 32767; i++) {
        const float a = (float)i, d = div32767(a), t = a / 32767.0f;
        if (memcmp(&d, &t, 4)) bad++;
    }
    printf("%d\n", bad);
    return bad != 0;
}
