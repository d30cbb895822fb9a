"""Pins the oracle (and the TF-checkpoint reader) to TensorFlow itself -- to be run ONCE on a box that has TensorFlow
2.3/2.4 (the reference's pinned version, docker/Dockerfile-amd64:9); neither TensorFlow nor a network exists in the build
container, which is why oracle/poisson_oracle.py says "parity unpinned" for the TF operators.

    python tests/golden/make_tf_fixtures.py [--reference /path/to/poisson_CNN/checkout]

writes into tests/golden/:
  tf_ops.npz        per-operator input/output pairs straight from TensorFlow: AveragePooling2D('same') on 200x300 and
                    109x130 for every stride the shipped config uses; tf.nn.conv2d_transpose('SAME', k == stride) for
                    strides 2,3,4,8,16; tf.image.resize nearest/bilinear/bicubic (up-sampling small maps, as
                    layers/Upsample.py:57 does) and the legacy tf.compat.v1.image.resize_bicubic(align_corners=True) of
                    dataset/utils/image_resize.py:20; tf.pad CONSTANT/SYMMETRIC/REFLECT; tf.nn.leaky_relu; inference
                    BatchNormalization; Conv2D(VALID) as cross-correlation; tf.linspace; tf.image.rot90.
  tf_forward.npz    (only with --reference) forward passes of the reference's own Homogeneous_Poisson_NN_Legacy,
                    Dirichlet_BC_NN_Legacy_2 and Poisson_CNN_Legacy on this repo's seeded synthetic weights and inputs; the
                    weights reach the Keras models through a checkpoint WRITTEN by poisson_cnn_b200.tf_checkpoint
                    (model.load_weights(...).assert_consumed() proves the key map), and a checkpoint WRITTEN by TensorFlow
                    (tf_ckpt/pcnn.index + .data-00000-of-00001) is stored for the reader test.
tests/test_tf_fixtures.py consumes whatever exists (skips otherwise); once the files are committed, section (c) of the
coverage table stops being "unpinned".
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def op_fixtures(tf):
    rng = np.random.RandomState(1234)
    fx = {}
    # --- AveragePooling2D(pool=s, strides=s, 'same'), channels_last inside TF (NCHW pooling has no CPU kernel)
    for (H, W) in ((200, 300), (109, 130)):
        x = rng.randn(1, H, W, 2).astype(np.float32)
        fx["pool_in_%dx%d" % (H, W)] = x
        for s in (2, 3, 4, 8, 16, 32, 64, 128):
            y = tf.keras.layers.AveragePooling2D(pool_size=s, strides=s, padding="same")(x).numpy()
            fx["pool_out_%dx%d_s%d" % (H, W, s)] = y
    # --- conv2d_transpose, SAME, k == stride (layers/deconvupscale.py:100-109); kernel [kh,kw,Cout,Cin]
    for s in (2, 3, 4, 8, 16):
        for (oh, ow) in ((37, 50), (64, 64)):
            ih, iw = -(-oh // s), -(-ow // s)
            x = rng.randn(2, ih, iw, 3).astype(np.float32)
            k = rng.randn(s, s, 4, 3).astype(np.float32)
            y = tf.nn.conv2d_transpose(x, k, [2, oh, ow, 4], strides=[1, s, s, 1], padding="SAME").numpy()
            fx["deconv_x_s%d_%dx%d" % (s, oh, ow)] = x
            fx["deconv_k_s%d_%dx%d" % (s, oh, ow)] = k
            fx["deconv_y_s%d_%dx%d" % (s, oh, ow)] = y
    # --- tf.image.resize (TF2 semantics: half-pixel centres, antialias=False)
    for (ih, iw, oh, ow) in ((2, 2, 256, 256), (4, 4, 200, 300), (8, 8, 109, 130), (7, 5, 64, 77)):
        x = rng.randn(1, ih, iw, 2).astype(np.float32)
        fx["resize_in_%dx%d" % (ih, iw)] = x
        for m in ("nearest", "bilinear", "bicubic"):
            fx["resize_%s_%dx%d_to_%dx%d" % (m, ih, iw, oh, ow)] = tf.image.resize(x, [oh, ow], method=m).numpy()
    # --- legacy bicubic with align_corners (dataset/utils/image_resize.py:20)
    for (ih, iw, oh, ow) in ((5, 5, 64, 64), (8, 3, 100, 37)):
        x = rng.randn(1, ih, iw, 1).astype(np.float32)
        fx["legacy_bicubic_in_%dx%d" % (ih, iw)] = x
        fx["legacy_bicubic_%dx%d_to_%dx%d" % (ih, iw, oh, ow)] = tf.compat.v1.image.resize_bicubic(x, [oh, ow], align_corners=True).numpy()
    # --- tf.pad
    x = rng.randn(1, 2, 6, 7).astype(np.float32)
    fx["pad_in"] = x
    for mode in ("CONSTANT", "SYMMETRIC", "REFLECT"):
        for k in (3, 4, 7):
            lo, hi = k // 2, k // 2 - (1 - k % 2)            # utils/apply_advanced_padding_and_call_conv_layer.py:8-14
            fx["pad_%s_k%d" % (mode, k)] = tf.pad(x, [[0, 0], [0, 0], [lo, hi], [lo, hi]], mode=mode, constant_values=0.0 if mode != "CONSTANT" else 2.0).numpy()
    # --- activations, BN, conv, linspace, rot90
    v = np.linspace(-3, 3, 31).astype(np.float32)
    fx["leaky_in"], fx["leaky_out"] = v, tf.nn.leaky_relu(v).numpy()
    bn = tf.keras.layers.BatchNormalization(axis=1)
    x = rng.randn(2, 3, 5, 4).astype(np.float32)
    bn(x, training=False)
    g, b, m, var = (rng.rand(3).astype(np.float32) + 0.5, rng.randn(3).astype(np.float32), rng.randn(3).astype(np.float32), rng.rand(3).astype(np.float32) + 0.5)
    bn.set_weights([g, b, m, var])
    fx["bn_in"], fx["bn_gamma"], fx["bn_beta"], fx["bn_mean"], fx["bn_var"] = x, g, b, m, var
    fx["bn_out"] = bn(x, training=False).numpy()
    x = rng.randn(1, 9, 11, 3).astype(np.float32)
    k = rng.randn(3, 5, 3, 2).astype(np.float32)
    fx["conv_in"], fx["conv_k"] = x, k
    fx["conv_out"] = tf.nn.conv2d(x, k, strides=1, padding="VALID").numpy()
    for n in (2, 7, 256):
        fx["linspace_%d" % n] = tf.linspace(0.0, 1.0, n).numpy()
    x = rng.randn(1, 4, 6, 1).astype(np.float32)
    fx["rot_in"] = x
    for kk in (1, 2, 3):
        fx["rot90_k%d" % kk] = tf.image.rot90(x, k=kk).numpy()
    return fx


def forward_fixtures(tf, reference_root):
    sys.path.insert(0, reference_root)
    import poisson_CNN as ref                                   # the UNMODIFIED reference package
    import torch
    from poisson_cnn_b200 import load_experiment, weights as W, tf_checkpoint as T
    from poisson_cnn_b200.synthetic import make_problem
    cfg = load_experiment("pcnn_end_to_end")
    hp_cfg, db_cfg = cfg["hpnn_model"], cfg["dbcnn_model"]
    hs, ds = W.hpnn_weight_specs(hp_cfg, "hpnn/"), W.dbcnn_weight_specs(db_cfg, "dbcnn/")
    w = W.synthetic_weights(({**hs[0], **ds[0]}, {**hs[1], **ds[1]}), seed=0)
    hp = ref.models.Homogeneous_Poisson_NN_Legacy(**ref.convert_tf_object_names(hp_cfg))
    db = ref.models.Dirichlet_BC_NN_Legacy_2(**ref.convert_tf_object_names(db_cfg))
    model = ref.models.Poisson_CNN_Legacy(hp, db)
    p = make_problem(1, 112, 120, seed=1003)                    # the inputs of tests/golden/pcnn_112x120.npz
    keys = ("rhs", "left", "top", "right", "bottom", "dx")
    inp = [tf.constant(p[k].numpy()) for k in keys]
    model(inp)                                                  # creates the variables
    ckdir = os.path.join(HERE, "tf_ckpt")
    os.makedirs(ckdir, exist_ok=True)
    ours = os.path.join(ckdir, "written_by_pcnn_b200")
    T.save_checkpoint_weights(ours, w, T.pcnn_key_map(hp_cfg, db_cfg))
    model.load_weights(ours).assert_existing_objects_matched()
    out = model(inp).numpy()
    model.save_weights(os.path.join(ckdir, "pcnn"))             # a TensorFlow-written checkpoint for the reader test
    fx = {k: p[k].numpy() for k in keys}
    fx["pcnn_out"] = out
    q = make_problem(2, 64, 64, seed=1001, magnitudes=False)
    hp_s = ref.models.Homogeneous_Poisson_NN_Legacy(**ref.convert_tf_object_names(load_experiment("hpnn_smalldomain")["model"]))
    fx["hpnn_rhs"], fx["hpnn_dx"] = q["rhs"].numpy(), q["dx"].numpy()
    fx["hpnn_out_shipped_cfg"] = hp([tf.constant(p["rhs"].numpy()), tf.constant(p["dx"].numpy())]).numpy()
    fx["dbcnn_out"] = db([tf.constant(p["left"].numpy()), tf.constant(p["dx"].numpy()), 112]).numpy()
    del hp_s
    return fx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=None, help="checkout of aligirayhanozbay/poisson_CNN (adds the model-level fixtures)")
    args = ap.parse_args()
    import tensorflow as tf
    print("TensorFlow", tf.__version__)
    np.savez_compressed(os.path.join(HERE, "tf_ops.npz"), tf_version=np.array(tf.__version__), **op_fixtures(tf))
    print("wrote tf_ops.npz")
    if args.reference:
        np.savez_compressed(os.path.join(HERE, "tf_forward.npz"), tf_version=np.array(tf.__version__), **forward_fixtures(tf, args.reference))
        print("wrote tf_forward.npz and tf_ckpt/")


if __name__ == "__main__":
    main()
