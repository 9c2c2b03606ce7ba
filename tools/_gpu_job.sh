for o in none density pca; do python tools/conn_probe.py --genes 8192 --perms 1184 --order $o --reps 2 | grep -o '"order": "[a-z]*"\|"kernel_ms": [^]]*]\|frac_of_8.2e7": [0-9.]*' | paste - - -; done
