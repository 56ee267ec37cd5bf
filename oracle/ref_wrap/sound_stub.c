/* oracle/ref_wrap/sound_stub.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Stand-in for the five audio back ends of the reference's _quisk extension (sound_alsa.c, sound_pulseaudio.c,
 * sound_portaudio.c, sound_directx.c, sound_wasapi.c) in an image that has none of their libraries.  Everything
 * else of _quisk (quisk.c, sound.c, microphone.c, ...) is compiled UNMODIFIED from /root/reference by
 * oracle/build_ref.sh step 4.  All entry points are no-ops except the "portaudio" playback, which records what
 * quisk_read_sound (sound.c:873) would have played -- the radio sound after quisk_process_samples -- so that
 * tests can read it back:  stub_capture_count(device_index), stub_capture_copy(device_index, out).
 */
#include <Python.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>
#include "quisk.h"

#define STUB_DEVS 8
static struct { complex double *buf; long n, cap; double volume; } cap_[STUB_DEVS];

static PyObject *two_empty_lists(void)
{
    PyObject *pylist = PyList_New(0), *a = PyList_New(0), *b = PyList_New(0);
    PyList_Append(pylist, a); PyList_Append(pylist, b);
    Py_DECREF(a); Py_DECREF(b);
    return pylist;
}
PyObject *quisk_alsa_sound_devices(PyObject *self, PyObject *args) { return two_empty_lists(); }
PyObject *quisk_directx_sound_devices(PyObject *self, PyObject *args) { return two_empty_lists(); }
PyObject *quisk_portaudio_sound_devices(PyObject *self, PyObject *args) { return two_empty_lists(); }
PyObject *quisk_pulseaudio_sound_devices(PyObject *self, PyObject *args) { return two_empty_lists(); }
PyObject *quisk_wasapi_sound_devices(PyObject *self, PyObject *args) { return two_empty_lists(); }
PyObject *quisk_alsa_control_midi(PyObject *self, PyObject *args, PyObject *kw) { Py_RETURN_NONE; }
PyObject *quisk_wasapi_control_midi(PyObject *self, PyObject *args, PyObject *kw) { Py_RETURN_NONE; }
void quisk_alsa_mixer_set(char *card, int numid, PyObject *value, char *err, int err_size) { if (err && err_size > 0) err[0] = 0; }

int quisk_read_alsa(struct sound_dev *dev, complex double *cs) { return 0; }
int quisk_read_portaudio(struct sound_dev *dev, complex double *cs) { return 0; }
int quisk_read_pulseaudio(struct sound_dev *dev, complex double *cs) { return 0; }
int quisk_read_directx(struct sound_dev *dev, complex double *cs) { return 0; }
int quisk_read_wasapi(struct sound_dev *dev, complex double *cs) { return 0; }
void quisk_play_alsa(struct sound_dev *dev, int n, complex double *cs, int report, double volume) {}
void quisk_play_pulseaudio(struct sound_dev *dev, int n, complex double *cs, int report, double volume) {}
void quisk_play_directx(struct sound_dev *dev, int n, complex double *cs, int report, double volume) {}
void quisk_play_wasapi(struct sound_dev *dev, int n, complex double *cs, double volume) {}
void quisk_write_wasapi(struct sound_dev *dev, int n, complex double *cs, double volume) {}
void quisk_alsa_sidetone(struct sound_dev *dev) {}
void quisk_pulseaudio_sidetone(struct sound_dev *dev) {}
void quisk_cork_pulseaudio(struct sound_dev *dev, int b) {}
void quisk_flush_pulseaudio(struct sound_dev *dev) {}
void quisk_start_sound_alsa(struct sound_dev **c, struct sound_dev **p) {}
void quisk_start_sound_pulseaudio(struct sound_dev **c, struct sound_dev **p) {}
void quisk_start_sound_directx(struct sound_dev **c, struct sound_dev **p) {}
void quisk_start_sound_wasapi(struct sound_dev **c, struct sound_dev **p) {}
void quisk_close_sound_alsa(struct sound_dev **c, struct sound_dev **p) {}
void quisk_close_sound_directx(struct sound_dev **c, struct sound_dev **p) {}
void quisk_close_sound_wasapi(struct sound_dev **c, struct sound_dev **p) {}
void quisk_close_sound_portaudio(void) {}
void quisk_close_sound_pulseaudio(void) {}

/* the recording "sound card": every playback device whose driver was set to DEV_DRIVER_PORTAUDIO */
void quisk_start_sound_portaudio(struct sound_dev **c, struct sound_dev **p)
{
    int i;
    for (i = 0; p && p[i]; i++)
        if (p[i]->driver == DEV_DRIVER_PORTAUDIO) {
            p[i]->handle = (void *)&cap_[i % STUB_DEVS];
            p[i]->rate_min = p[i]->rate_max = p[i]->sample_rate;
            p[i]->chan_min = p[i]->chan_max = 2;
        }
}
void quisk_play_portaudio(struct sound_dev *dev, int n, complex double *cs, int report, double volume)
{
    int k = ((char *)dev->handle - (char *)cap_) / (int)sizeof(cap_[0]);
    if (!dev->handle || k < 0 || k >= STUB_DEVS || n <= 0 || !cs) return;
    if (cap_[k].n + n > cap_[k].cap) {
        cap_[k].cap = (cap_[k].n + n) * 2 + 4096;
        cap_[k].buf = (complex double *)realloc(cap_[k].buf, cap_[k].cap * sizeof(complex double));
    }
    memcpy(cap_[k].buf + cap_[k].n, cs, n * sizeof(complex double));
    cap_[k].n += n; cap_[k].volume = volume;
}
long stub_capture_count(int k) { return (k >= 0 && k < STUB_DEVS) ? cap_[k].n : 0; }
double stub_capture_volume(int k) { return (k >= 0 && k < STUB_DEVS) ? cap_[k].volume : 0.0; }
long stub_capture_copy(int k, complex double *out)
{
    if (k < 0 || k >= STUB_DEVS) return 0;
    memcpy(out, cap_[k].buf, cap_[k].n * sizeof(complex double));
    return cap_[k].n;
}
void stub_capture_clear(int k) { if (k >= 0 && k < STUB_DEVS) cap_[k].n = 0; }
