#!/bin/bash
# Is a real TensorFlow available on the GPU box to pin the oracle's third-party kernels (NonMaxSuppressionV5,
# tf.linalg.inv, tfp Categorical) against?  Records the answer under gpurun_out/ (copied to profiles/).
out=gpurun_out/tf_probe_r2.txt
{
  echo "# TensorFlow probe on the GPU box, $(date -u +%Y-%m-%dT%H:%M:%SZ)"
  echo "## python -c 'import tensorflow'"
  python -c "import tensorflow as tf; print('tensorflow', tf.__version__)" 2>&1 | tail -3
  echo "## python -c 'import tensorflow_probability'"
  python -c "import tensorflow_probability as tfp; print('tfp', tfp.__version__)" 2>&1 | tail -2
  echo "## pip download tensorflow (index reachability, 20 s limit)"
  timeout 20 python -m pip download --no-deps -d /tmp/tfwheel tensorflow 2>&1 | tail -3
  echo "rc=$?"
  echo "## offline wheelhouse"
  ls /opt/wheelhouse 2>/dev/null | grep -i -E "tensorflow|tf_|keras" || echo "no tensorflow wheel under /opt/wheelhouse"
  echo "## site-packages"
  python -m pip list 2>/dev/null | grep -i -E "tensorflow|keras|jax" || echo "no tensorflow / keras / jax distribution installed"
} > $out 2>&1
cat $out
