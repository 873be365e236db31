#!/bin/bash
# Is there any way to compile the Fortran reference on this box?  (VERDICT r01 next-round 1c,
# SURVEY.md 8c item 3.)  Output is committed under profiles/ as the record of the probe.
echo "# Fortran toolchain probe on $(hostname) at $(date -u +%FT%TZ)"
nvidia-smi --query-gpu=name --format=csv,noheader 2>/dev/null | head -1
for c in gfortran gfortran-13 gfortran-12 gfortran-11 flang flang-new lfortran ifort ifx f77 f95 f2c nvfortran pgfortran g77 fort77; do
  p=$(command -v $c 2>/dev/null); echo "$c: ${p:-absent}"
done
echo "f951 (the gfortran compiler proper):"; find / -name 'f951*' -not -path '/proc/*' 2>/dev/null | head; echo "(end)"
echo "libgfortran runtimes:"; find / -name 'libgfortran*' -not -path '/proc/*' 2>/dev/null | head; echo "(end)"
echo "netcdf:"; find / \( -name 'libnetcdff*' -o -name 'netcdf.mod' -o -name 'netcdf.inc' \) -not -path '/proc/*' 2>/dev/null | head; echo "(end)"
echo "gcc: $(gcc --version | head -1)"; echo "languages: $(gcc -v 2>&1 | grep -o 'enable-languages=[^ ]*')"
python -c "import numpy.f2py, numpy; print('numpy.f2py importable, numpy', numpy.__version__)" 2>&1 | tail -1
