/* oracle/ref_wrap/quisk_tx_wrap.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Turns the `static` transmit-audio functions of the reference's microphone.c into a loadable library WITHOUT copying
 * them into this repository: oracle/build_ref.sh extracts the line ranges below from /root/reference/microphone.c into a
 * scratch file at build time and this wrapper `#include`s it.  Everything in THIS file is our own glue: the file-scope
 * variables those functions read (microphone.c:29-37, 65; quisk.c:111) and `ref_tx_*` accessors for ctypes.
 *
 * Extracted ranges (microphone.c): 42-56 struct alc, 161-233 CcmPeak, 235-370 init_alc / process_alc, 372-604 tx_filter,
 * 605-624 tx_filter_digital.
 */
#include <Python.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <complex.h>
#include "quisk.h"
#include "filter.h"

#define DEBUG 0
#define DEBUG_IO 0
#define MIC_OUT_RATE 48000              /* microphone.c:29-31 */

struct sound_conf quisk_sound_state;
rx_mode_type rxMode;
double quisk_mic_preemphasis;
double quisk_mic_clip;
static double mic_agc_level = 0.10;     /* microphone.c:65 */

#define DEBUG_LEVEL 0
#include "quisk_tx_funcs.inc"           /* microphone.c:42-56, 161-233, 235-370, 372-624 */

void ref_tx_init(int mode, int mic_sample_rate, double preemphasis, double clip)
{
    rxMode = (rx_mode_type)mode;
    quisk_sound_state.mic_sample_rate = mic_sample_rate;
    quisk_mic_preemphasis = preemphasis;
    quisk_mic_clip = clip;
    tx_filter(NULL, 0);
}

int ref_tx_filter(complex double *samples, int count) { return tx_filter(samples, count); }

void ref_tx_digital_init(int mode) { rxMode = (rx_mode_type)mode; tx_filter_digital(NULL, 0); }
int ref_tx_filter_digital(complex double *samples, int count) { return tx_filter_digital(samples, count); }

static struct alc tx_alc;
void ref_alc_init(void) { init_alc(&tx_alc, 960); init_alc(&tx_alc, 0); }      /* microphone.c:1178, 1207 */
void ref_alc_key_down(void) { init_alc(&tx_alc, 0); }
void ref_process_alc(complex double *samples, int count, int mode) { process_alc(samples, count, &tx_alc, (rx_mode_type)mode); }
