# development only: build an experimental variant of the library with extra -D flags, next to the product build.
#   bash tools/build_variant.sh r8 "-DDD_K4_R_LARGE=8"   ->  distdiff_b200/_variants/r8.so   (load it with DD_LIB_PATH=...)
name=$1; shift
mkdir -p distdiff_b200/_variants
DD_EXTRA_NVCC_FLAGS="$*" DD_BUILD_DIR=$PWD/distdiff_b200/_variants/_build_$name DD_LIB_OUT=$PWD/distdiff_b200/_variants/$name.so python -m distdiff_b200.build
