#!/bin/bash
# Regenerates profiles/r02_sass_evidence.txt from the production objects (run after `make -C voice-activity-detection-vad-onnx_b200/csrc`).
B=voice-activity-detection-vad-onnx_b200/csrc/build
echo "# SASS evidence (cuobjdump -sass of the sm_100a objects that make libvadx.so; counts per object, then excerpts)"
echo "# UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UBLKCP.S.G / UBLKCP.G.S = cp.async.bulk global->shared / shared->global,"
echo "# UTMASTG = cp.async.bulk.tensor store (2-D tensor map), UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier, FFMA2 = fma.rn.f32x2,"
echo "# USETMAXREG = setmaxnreg; legacy_HMMA = mma.sync (none: every match of 'HMMA' is the UTCHMMA mnemonic)"
for o in gemm_tc stft_tc block_stages memory_bulk elementwise iccrn silero_extra; do
  cuobjdump -sass $B/$o.o > /tmp/_s.txt
  c() { grep -c "$1" /tmp/_s.txt; }
  printf "%-14s UTCHMMA=%s LDTM=%s STTM=%s UBLKCP.S.G=%s UBLKCP.G.S=%s UTMASTG=%s UBLKPF=%s UTMALDG=%s SYNCS=%s FFMA2=%s USETMAXREG=%s legacy_HMMA=%s LDGSTS=%s\n" $o \
    $(c UTCHMMA) $(c "LDTM") $(c STTM) $(c "UBLKCP.S.G") $(c "UBLKCP.G.S") $(c UTMASTG) $(c UBLKPF) $(c UTMALDG) $(c SYNCS) $(c FFMA2) $(c USETMAXREG) \
    $(grep "HMMA" /tmp/_s.txt | grep -vc UTCHMMA) $(c LDGSTS)
done
echo
echo "## gemm_tc.o: first tcgen05.mma / tcgen05.ld / bulk-copy / tensor-map store instructions"
cuobjdump -sass $B/gemm_tc.o > /tmp/_s.txt
for k in "UTCHMMA" "LDTM" "UBLKCP.S.G" "UBLKCP.G.S" "UTMASTG" "UBLKPF"; do grep -m 3 "$k" /tmp/_s.txt; done
echo "## block_stages.o (fc2_memory_stages_kernel): transposed product, TMEM window loads, packed FMAs, register hand-over, bulk copies"
cuobjdump -sass $B/block_stages.o > /tmp/_s.txt
for k in "UTCHMMA" "LDTM" "FFMA2" "USETMAXREG" "UBLKCP.S.G"; do grep -m 3 "$k" /tmp/_s.txt; done
echo "## stft_tc.o (stft_power_tc_kernel)"
cuobjdump -sass $B/stft_tc.o > /tmp/_s.txt
for k in "UTCHMMA" "LDTM" "UBLKCP"; do grep -m 2 "$k" /tmp/_s.txt; done
echo "## memory_bulk.o (fsmn_memory_bulk_kernel)"
cuobjdump -sass $B/memory_bulk.o > /tmp/_s.txt
for k in "FFMA2" "UBLKCP"; do grep -m 2 "$k" /tmp/_s.txt; done
