#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference (GenomicsDB v0.10.2 + vendored TileDB + htslib)
# from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored, travels with gpurun).
# Nothing is copied into the repo; the reference's own build system (cmake) is not used because it insists on
# MPI / libuuid / protobuf (CMakeLists.txt:92-97,120,150).  Shims under oracle/ref_shim/ replace those three
# (none touches the combine path).  Recipe follows SURVEY.md Appendix B.
#
# Outputs: oracle/_ref/bin/{gt_mpi_gather,vcf2tiledb,create_tiledb_workspace,consolidate_tiledb_array,gdb_ref_dump}
#          oracle/_ref/lib/libgdbref.a (all reference objects, for the harness)
set -u
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${GDB_REFERENCE:-/root/reference}"
W="$HERE/_ref"
SH="$HERE/ref_shim"
J="${JOBS:-$(nproc)}"
if [ ! -d "$R/src/main/cpp" ]; then echo "reference not present at $R: keeping prebuilt oracle/_ref"; exit 0; fi
mkdir -p "$W/obj_h/cram" "$W/obj_t" "$W/obj_g" "$W/bin" "$W/lib"
STAMP="$W/.built_ok"
if [ -f "$STAMP" ] && [ "${FORCE:-0}" != "1" ] && [ "$HERE/ref_harness/gdb_ref_dump.cc" -ot "$STAMP" ] && [ "$0" -ot "$STAMP" ]; then
  echo "oracle/_ref up to date"; exit 0; fi
rm -f "$STAMP"
fail=0
H="$R/dependencies/htslib"
HOBJS="kfunc knetfile kstring bcf_sr_sort bgzf errmod faidx hfile hfile_net hts hts_os md5 multipart probaln realn regidx sam synced_bcf_reader vcf_sweep tbx textutils thread_pool vcf vcfutils cram/cram_codecs cram/cram_decode cram/cram_encode cram/cram_external cram/cram_index cram/cram_io cram/cram_samtools cram/cram_stats cram/files cram/mFILE cram/open_trace_file cram/pooled_alloc cram/rANS_static cram/sam_header cram/string_alloc"
echo $HOBJS | tr ' ' '\n' | xargs -P"$J" -I{} sh -c "[ -f $W/obj_h/{}.o ] || gcc -O2 -fPIC -w -I$SH/htslib_gen -I$H -c $H/{}.c -o $W/obj_h/{}.o || echo FAIL {}" | grep FAIL && fail=1
TINC=$(for d in "$R"/dependencies/TileDB/core/include/*; do echo -n "-I$d "; done)
CXXF="-std=c++11 -O3 -DNDEBUG -w -fPIC -D_FILE_OFFSET_BITS=64 -I$SH -include $SH/prefix.h"
ls "$R"/dependencies/TileDB/core/src/*/*.cc | grep -v -E "storage_hdfs|storage_gcs" | \
  xargs -P"$J" -I{} sh -c 'b=$(basename {} .cc); [ -f '"$W"'/obj_t/$b.o ] || g++ '"$CXXF $TINC"' -c {} -o '"$W"'/obj_t/$b.o || echo FAIL {}' | grep FAIL && fail=1
gcc -O2 -fPIC -c "$SH/uuid_shim.c" -o "$W/obj_t/uuid_shim.o" || fail=1
C="$R/src/main/cpp"
GINC="-I$C/include/genomicsdb -I$C/include/loader -I$C/include/query_operations -I$C/include/utils -I$C/include/vcf -I$C/include/config -I$R/dependencies/RapidJSON/include -I$H -I$R/dependencies/TileDB/core/include/c_api"
GDEF="-DHTSDIR=1 -DDUPLICATE_CELL_AT_END=1 -DDISABLE_OPENMP=1 ${EXTRA_DEFS:-}"
ls "$C"/src/*/*.cc | grep -v vid_mapper_pb | \
  xargs -P"$J" -I{} sh -c 'b=$(basename {} .cc); [ -f '"$W"'/obj_g/$b.o ] || g++ '"$CXXF $GDEF $GINC $TINC"' -c {} -o '"$W"'/obj_g/$b.o || echo FAIL {}' | grep FAIL && fail=1
g++ $CXXF $GDEF $GINC $TINC -c "$SH/pb_stub.cc" -o "$W/obj_g/pb_stub.o" || fail=1
rm -f "$W/lib/libgdbref.a"
ar rcs "$W/lib/libgdbref.a" "$W"/obj_g/*.o "$W"/obj_t/*.o $(for o in $HOBJS; do echo "$W/obj_h/$o.o"; done) || fail=1
for t in gt_mpi_gather vcf2tiledb create_tiledb_workspace consolidate_tiledb_array; do
  g++ $CXXF $GDEF $GINC $TINC -I"$R/tools/include" "$R/tools/src/$t.cc" -Wl,--start-group "$W/lib/libgdbref.a" -Wl,--end-group -lz -lcrypto -lpthread -o "$W/bin/$t" || { echo LINKFAIL $t; fail=1; }
done
# our harness (dump operator + null-adapter timing) on top of the unmodified reference objects
g++ $CXXF $GDEF $GINC $TINC -I"$R/tools/include" "$HERE/ref_harness/gdb_ref_dump.cc" -Wl,--start-group "$W/lib/libgdbref.a" -Wl,--end-group -lz -lcrypto -lpthread -o "$W/bin/gdb_ref_dump" || { echo LINKFAIL gdb_ref_dump; fail=1; }
if [ $fail = 0 ]; then touch "$STAMP"; echo BUILD_OK; else echo BUILD_FAILED; exit 1; fi
