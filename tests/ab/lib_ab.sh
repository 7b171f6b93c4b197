#!/bin/bash
# alternate two builds of the library on the same box: tests/ab/lib_ab.sh B N k rounds
B=$1; N=$2; K=$3; R=${4:-3}
PREV=crowd-coachable-recommendations_b200/lib/libccr_b200_prev.so
for i in $(seq $R); do
  echo -n "prev "; CCR_B200_LIB=$PREV python tests/ncu_case.py $B $N $K 0 8 | tail -1
  echo -n "curr "; python tests/ncu_case.py $B $N $K 0 8 | tail -1
done
