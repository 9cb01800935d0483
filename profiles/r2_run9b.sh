mkdir -p gpurun_out
python profiles/probes/py_second_create.py 1000 > gpurun_out/r2i_second_create.txt 2>&1; grep -v "spgemm:" gpurun_out/r2i_second_create.txt | head -120
