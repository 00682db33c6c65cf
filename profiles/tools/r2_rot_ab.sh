#!/bin/bash
for lib in "$@"; do echo "== $lib"; for n in 100 200; do QIO_B200_LIB=$PWD/$lib python profiles/tools/rotate_profile.py $n 3; done; done
