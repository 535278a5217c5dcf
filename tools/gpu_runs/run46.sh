cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
for i in 1 2; do
echo "main: $(timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1)"
echo "erel: $(SCIPNP_LIB=$E/libscipnp_erel.so timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1)"
done
echo "erel 278: $(SCIPNP_LIB=$E/libscipnp_erel.so timeout 200 python profiles/prof_driver.py 40 278 3840 24 2>&1 | tail -1)"
echo "main 278: $(timeout 200 python profiles/prof_driver.py 40 278 3840 24 2>&1 | tail -1)"
