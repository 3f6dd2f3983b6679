#!/usr/bin/env bash
# TEST/BENCH INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (OHDSI/bayes-bridge v0.2.6)
# into oracle/_ref/ so that tests and bench.py's cpu baseline can run the reference's own
# numpy/scipy/Cython sampler on the same host.  Nothing under bayesbridge_b200/ imports it.
#
# The reference tree is read-only, so it is copied to a scratch dir first.  Two one-token
# compatibility shims are applied to the COPY (never to /root/reference), both required by
# the scipy (1.18) in this image and neither changing arithmetic:
#   (i)  reg_coef_sampler/cg_sampler.py:78   cg(..., tol=rtol)  ->  cg(..., rtol=rtol)
#        (scipy >= 1.14 removed `tol`; same stopping rule ||r|| < rtol*||b||)
#   (ii) reg_coef_sampler/direct_gaussian_sampler.py:22  cholesky(A, scale_array) -> cholesky(A, lower=False)
#        (the array was silently read as `lower`; upper factor matches cho_solve((chol, False)) on :23-25)
# oracle/_ref/ is git-ignored (reference sources never enter history) but ships to the GPU box.
set -euo pipefail
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF absent (GPU box?) -- using prebuilt $OUT if present"; exit 0
fi
TMP=$(mktemp -d /tmp/bbref.XXXXXX)
cp -r "$REF"/. "$TMP"/
rm -rf "$TMP/.git"
sed -i 's/maxiter=maxiter, tol=rtol,/maxiter=maxiter, rtol=rtol,/' "$TMP/bayesbridge/reg_coef_sampler/cg_sampler.py"
sed -i 's/sp.linalg.cholesky(Prec_precond, jacobi_precond_scale)/sp.linalg.cholesky(Prec_precond, lower=False)/' \
    "$TMP/bayesbridge/reg_coef_sampler/direct_gaussian_sampler.py"
grep -q 'rtol=rtol' "$TMP/bayesbridge/reg_coef_sampler/cg_sampler.py"
grep -q 'lower=False)' "$TMP/bayesbridge/reg_coef_sampler/direct_gaussian_sampler.py"
( cd "$TMP" && python setup.py -q build_ext --inplace >"$TMP/build.log" 2>&1 ) || { tail -30 "$TMP/build.log"; exit 1; }
rm -rf "$OUT"; mkdir -p "$OUT"
cp -r "$TMP/bayesbridge" "$OUT/bayesbridge"
cp "$TMP/simulate_data.py" "$OUT/simulate_data.py"
mkdir -p "$OUT/tests_saved_outputs"
cp "$TMP"/tests/regression_tests/saved_outputs/*.npy "$OUT/tests_saved_outputs/"
find "$OUT" -name '*.c' -newer "$TMP/setup.py" -path '*random*' ! -name 'scipy_ndtr.c' -delete || true
find "$OUT" -name '*.ipynb' -delete; rm -rf "$OUT"/bayesbridge/**/build
rm -rf "$TMP"
echo "build_ref: reference built into $OUT"
