#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE, not product code.
#
# Builds the reference's own C implementation of the hot path from the sources
# where they lie under /root/reference into shared libraries under oracle/_ref/
# (git-ignored; they travel to the GPU box with the snapshot).  Nothing from
# /root/reference is copied into the repository: scratch files live in a
# mktemp directory that is removed on exit.
#
#   _ref/libquisk_filter_ref.so  filter.c, verbatim (all 17 filter.h functions + filters.h tables)
#   _ref/libquisk_rx_ref.so      filter.c + the static RX functions of quisk.c (see ref_wrap/quisk_rx_wrap.c)
#   _ref/libquisk_rx_ref_O3.so   the same at -O3: what bench.py times as the CPU reference
#   _ref/libquisk_rx_dropin.so   the same RX functions of quisk.c linked against quisk_b200/libquisk_cuda.so instead of filter.c
#   _ref/libquisk_tx_ref.so      filter.c + tx_filter / CcmPeak of microphone.c (see ref_wrap/quisk_tx_wrap.c)
#   _ref/libquisk_wdspglue_ref.so  wdspFexchange0 of quisk_wdsp.c (the re-blocker in front of fexchange0)
#   _ref/libwdsp_ref.so          wdsp/*.c (minus the make_*.c table generators) + our FFTW-API shim
#
# Flags: -O2 for the Quisk sources (setuptools default), -O3 for WDSP
# (wdsp/Makefile:9); never -ffast-math.
set -euo pipefail
REF=${QUISK_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
    echo "build_ref: $REF not present; keeping prebuilt files in $OUT" >&2
    exit 0
fi
PYINC=$(python3 -c 'import sysconfig; print(sysconfig.get_paths()["include"])')
if [ ! -f "$PYINC/Python.h" ]; then PYINC=/usr/include/python3.12; fi
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT

# 1. filter.c verbatim
gcc -O2 -fPIC -shared -w -I"$PYINC" -I"$REF" "$REF/filter.c" -o "$OUT/libquisk_filter_ref.so" -lm

# 2. quisk.c RX functions through the wrapper TU
{ sed -n '46,53p' "$REF/quisk.c"; sed -n '68,81p' "$REF/quisk.c"; } > "$TMP/quisk_rx_consts.inc"
{ sed -n '622,665p' "$REF/quisk.c"; sed -n '1182,1256p' "$REF/quisk.c"; sed -n '1633,1671p' "$REF/quisk.c";
  sed -n '1673,1846p' "$REF/quisk.c"; sed -n '1848,2160p' "$REF/quisk.c"; sed -n '2162,2287p' "$REF/quisk.c"; } > "$TMP/quisk_rx_funcs.inc"
{ sed -n '1056,1084p' "$REF/quisk.c"; sed -n '1086,1180p' "$REF/quisk.c"; } > "$TMP/quisk_squelch.inc"     # d_delay, ssb_squelch
sed -n '786,963p' "$REF/quisk.c" > "$TMP/quisk_notch.inc"                 # dAutoNotch
sed -n '679,784p' "$REF/quisk.c" > "$TMP/quisk_nb.inc"                    # NoiseBlanker (optional stage in front of the path)
sed -n '2922,2953p' "$REF/quisk.c" > "$TMP/quisk_unpack_py.inc"          # add_rx_samples: the two unpack branches
sed -n '3746,3763p' "$REF/quisk.c" > "$TMP/quisk_unpack_hermes.inc"      # read_rx_udp10: 24-bit record loop of one 512-byte frame
gcc -O2 -fPIC -shared -w -I"$PYINC" -I"$REF" -I"$TMP" -I"$HERE/fftw_shim" "$HERE/ref_wrap/quisk_rx_wrap.c" "$REF/filter.c" \
    "$HERE/fftw_shim/fftw_shim.c" "$HERE/fft64.c" -o "$OUT/libquisk_rx_ref.so" -lm

# 2a. The timing arm of bench.py uses an -O3 build of the same sources (the fixtures stay on the
#     setuptools default -O2 above -- without -ffast-math the two produce the same doubles, tests/test_oracle_vs_ref.py checks it)
gcc -O3 -fPIC -shared -w -I"$PYINC" -I"$REF" -I"$TMP" -I"$HERE/fftw_shim" "$HERE/ref_wrap/quisk_rx_wrap.c" "$REF/filter.c" \
    "$HERE/fftw_shim/fftw_shim.c" "$HERE/fft64.c" -o "$OUT/libquisk_rx_ref_O3.so" -lm

# 2b. The same wrapper TU linked against libquisk_cuda.so INSTEAD of filter.c: the reference's own orchestrator code
#     (quisk_process_decimate / quisk_process_demodulate, unmodified) calling the GPU filter.h drop-in.  This is the
#     swap-in of INTEGRATION.md section 1 in miniature; filters.h (the coefficient tables filter.c used to emit) is
#     compiled from the reference's header through a two-line scratch TU.  tests/test_dropin_gpu.py runs it.
CUDALIB="$HERE/../quisk_b200/libquisk_cuda.so"
if [ -f "$CUDALIB" ]; then
    printf '#include <complex.h>\n#include "filters.h"\n' > "$TMP/filters_data.c"
    gcc -O2 -fPIC -shared -w -I"$PYINC" -I"$REF" -I"$TMP" -I"$HERE/fftw_shim" "$HERE/ref_wrap/quisk_rx_wrap.c" "$TMP/filters_data.c" \
        "$HERE/fftw_shim/fftw_shim.c" "$HERE/fft64.c" -o "$OUT/libquisk_rx_dropin.so" -L"$HERE/../quisk_b200" -lquisk_cuda -Wl,-rpath,'$ORIGIN/../../quisk_b200' -lm
fi

# 2c. Quisk's side of the WDSP boundary: wdspFexchange0 (quisk_wdsp.c:7-69) through its own wrapper TU
sed -n '7,69p' "$REF/quisk_wdsp.c" > "$TMP/quisk_wdsp_glue.inc"
gcc -O2 -fPIC -shared -w -I"$TMP" "$HERE/ref_wrap/quisk_wdsp_wrap.c" -o "$OUT/libquisk_wdspglue_ref.so" -lm

# 2d. The transmit-audio chain of microphone.c (tx_filter + CcmPeak + tx_filter_digital + process_alc) through its own wrapper TU: the TX mirror of the
#     receive path, same filter.c underneath
{ sed -n '42,56p' "$REF/microphone.c"; sed -n '161,233p' "$REF/microphone.c"; sed -n '235,370p' "$REF/microphone.c"; sed -n '372,624p' "$REF/microphone.c"; } > "$TMP/quisk_tx_funcs.inc"
gcc -O2 -fPIC -shared -w -I"$PYINC" -I"$REF" -I"$TMP" "$HERE/ref_wrap/quisk_tx_wrap.c" "$REF/filter.c" -o "$OUT/libquisk_tx_ref.so" -lm

# 3. WDSP against the FFTW shim
WSRC=$(ls "$REF"/wdsp/*.c | grep -v '/make_')
mkdir -p "$TMP/wobj"
i=0
for f in $WSRC; do
    gcc -O3 -fPIC -w -D_GNU_SOURCE -I"$HERE/fftw_shim" -I"$REF/wdsp" -c "$f" -o "$TMP/wobj/$(basename "$f" .c).o" &
    i=$((i+1)); if [ $((i % 8)) -eq 0 ]; then wait; fi
done
wait
gcc -O3 -fPIC -w -c "$HERE/fftw_shim/fftw_shim.c" -o "$TMP/wobj/_fftw_shim.o"
gcc -O3 -fPIC -w -c "$HERE/fft64.c" -o "$TMP/wobj/_fft64.o"
gcc -shared -o "$OUT/libwdsp_ref.so" "$TMP"/wobj/*.o -lm -lpthread
# 4. The reference's WHOLE _quisk extension (setup.py:17-20 source list; the five audio back ends replaced by
#    ref_wrap/sound_stub.c, FFTW3 by the shim), twice: against its own filter.c, and against libquisk_cuda.so instead
#    (the filter.h swap-in of INTEGRATION.md section 1 under the real quisk_process_samples / quisk_read_sound), plus
#    the B4 sample-source plugin quisk_b200/plugin/quisk_block_source.c built like the reference's hardware plugins.
QSRC="quisk.c sound.c is_key_down.c microphone.c utility.c extdemod.c freedv.c quisk_wdsp.c ac2yd/remote.c tci.c base64.c handshake.c sha1.c utf8.c ws.c"
mkdir -p "$TMP/qobj" "$OUT/quisk_full/ref" "$OUT/quisk_full/cuda"
for f in $QSRC; do
    gcc -O2 -fPIC -w -I"$PYINC" -I"$REF" -I"$HERE/fftw_shim" -c "$REF/$f" -o "$TMP/qobj/$(basename "$f" .c).o" &
done
gcc -O2 -fPIC -w -I"$PYINC" -I"$REF" -c "$REF/filter.c" -o "$TMP/filter.o" &
gcc -O2 -fPIC -w -I"$PYINC" -I"$REF" -c "$HERE/ref_wrap/sound_stub.c" -o "$TMP/qobj/_sound_stub.o" &
gcc -O2 -fPIC -w -c "$HERE/fftw_shim/fftw_shim.c" -o "$TMP/qobj/_fftw_shim.o" &
gcc -O2 -fPIC -w -c "$HERE/fft64.c" -o "$TMP/qobj/_fft64.o" &
wait
gcc -shared -o "$OUT/quisk_full/ref/_quisk.so" "$TMP"/qobj/*.o "$TMP/filter.o" -lm -lpthread
if [ -f "$CUDALIB" ]; then
    gcc -O2 -fPIC -w -I"$PYINC" -I"$REF" -c "$TMP/filters_data.c" -o "$TMP/filters_data.o"
    gcc -shared -o "$OUT/quisk_full/cuda/_quisk.so" "$TMP"/qobj/*.o "$TMP/filters_data.o" \
        -L"$HERE/../quisk_b200" -lquisk_cuda -Wl,-rpath,'$ORIGIN/../../../../quisk_b200' -lm -lpthread
fi
gcc -O2 -fPIC -shared -w -I"$PYINC" -I"$REF" "$HERE/../quisk_b200/plugin/quisk_block_source.c" "$REF/import_quisk_api.c" \
    -o "$OUT/quisk_full/quisk_block_source.so"
echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')"
