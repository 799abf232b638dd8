#!/bin/bash
# builds tools/study/librt_study.so against oracle/librt_oracle.so (make -C oracle port first)
cd "$(dirname "$0")" && gcc -O2 -std=c99 -ffp-contract=off -fPIC -shared -Wall -Wextra -I../../oracle -o librt_study.so rt_study.c \
    -L../../oracle -lrt_oracle -Wl,-rpath,'$ORIGIN/../../oracle'
