/* oracle/ref_wrap/quisk_wdsp_wrap.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Quisk's side of the WDSP boundary, wdspFexchange0 (quisk_wdsp.c:22-69), as a loadable library WITHOUT copying it
 * into this repository: oracle/build_ref.sh extracts quisk_wdsp.c:7-69 (the CLIP32 constants, the static channel
 * table, the fexchange0 function pointer and wdspFexchange0 itself) into a scratch file at build time and this
 * wrapper #includes it.  The rest of that file is Python argument parsing (quisk_wdsp_set_parameter, :71-91); the
 * three assignments it makes are restated below as plain C setters for ctypes.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <complex.h>

#include "quisk_wdsp_glue.inc"      /* quisk_wdsp.c:7-69 */

void ref_wdsp_set_parameter(int channel, int in_size, void *fexchange0, int in_use)
{   /* quisk_wdsp.c:80-87 */
    if (channel >= 0 && channel < MAX_CHANNELS) {
        if (fexchange0)
            wdsp_fexchange0 = fexchange0;
        if (in_size > 0)
            wdspChannel[channel].in_size = in_size;
        if (in_use >= 0)
            wdspChannel[channel].in_use = in_use;
    }
}
