#!/bin/bash
# Does the synchronous batched learner learn a tracker?  Trains Track2D-BlockPartialPZR-v0 (full AD-VAT) and evaluates the tracker
# with the README protocol (gym_eval on Block/Maze x Nav/Ram, 100 greedy episodes).  usage: tools/learn_experiment.sh TAG E ITERS EVAL_EVERY [extra train args]
cd ${GRAFT_REPO_ROOT:-.}
TAG=$1; E=$2; ITERS=$3; EVERY=$4; shift 4
OUT=gpurun_out/learn_$TAG; LOG=/tmp/learn_$TAG
mkdir -p $OUT $LOG
python -m active_tracking_rl_b200.train --env Track2D-BlockPartialPZR-v0 --num-envs $E --iters $ITERS --eval-every $EVERY --graph --split \
    --max-step 100000000 --log-dir $LOG/ "$@" > $OUT/train.log 2>&1
D=$(ls -d $LOG/*/* | head -1)
cp $D/tracker-best.dat $D/target-best.dat $OUT/ 2>/dev/null
cp $D/tracker-new.dat $OUT/ 2>/dev/null
grep "^eval" $OUT/train.log > $OUT/evals.txt
for ENV in Track2D-BlockPartialNav-v0 Track2D-BlockPartialRam-v0 Track2D-MazePartialNav-v0 Track2D-MazePartialRam-v0; do
  for W in best new; do
    [ -f $D/tracker-$W.dat ] && echo "$ENV tracker-$W: $(python -m active_tracking_rl_b200.gym_eval --env $ENV --network tat-maze-lstm --load-tracker $D/tracker-$W.dat --num-episodes 100 --csv $OUT/eval_$W.csv 2>&1 | tail -1)" >> $OUT/final_eval.txt
  done
done
tail -3 $OUT/train.log; cat $OUT/evals.txt | tail -20; cat $OUT/final_eval.txt
