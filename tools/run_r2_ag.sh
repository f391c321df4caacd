# Round-2 GPU call AG: window size of a round of 8 x 2^20 commitments (plan follows the largest job: c = 15)
mkdir -p gpurun_out
( python tools/time_batch.py 20; SNARKVM_B200_MSM_C=16 python tools/time_batch.py 20; SNARKVM_B200_MSM_C=16 SNARKVM_B200_MSM_LEVELS=4 python tools/time_batch.py 20; SNARKVM_B200_MSM_C=17 SNARKVM_B200_MSM_LEVELS=4 python tools/time_batch.py 20; SNARKVM_B200_MSM_C=15 SNARKVM_B200_MSM_LEVELS=4 python tools/time_batch.py 20 ) > gpurun_out/r2ag_batch.log 2>&1
cat gpurun_out/r2ag_batch.log
