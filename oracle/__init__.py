"""CPU oracle for the detection hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain restatement (torch-CPU fp32 convolutions, numpy fp32 arithmetic, a C
NonMaxSuppressionV3) of the inference graphs the reference builds with
TensorFlow 1.13 ops (SSD300.py, SSD512.py, RetinaNet.py, YOLOv3.py, FCOS.py of
Stick-To/Object-Detection-Tensorflow).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package;
the product package (object-detection-tensorflow_b200/odt_b200) never does.

PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors for
this path, TensorFlow 1.13 is not installable here, and
SSD300.py / SSD512.py do not even parse (empty `else:` at line 41-43).  The
TF-op semantics restated in oracle/tfops.py (SAME padding, BN eps 1e-3,
GroupNorm eps 1e-6, legacy bilinear resize, NonMaxSuppressionV3 ...) come from
knowledge of TF 1.13's published kernels and could not
be checked against a running TensorFlow.  The pins are the hand-derived known
answers of SURVEY.md section 8(c) (anchor tables) and, for the ops whose definition
does not depend on TensorFlow, independent implementations (torchvision.ops.nms for
the greedy strict-> suppression on tie-free boxes, torch.nn.functional for group
norm / softmax / l2-normalise / max-pool / nearest resize) -- tests/test_oracle.py.

oracle/loss.py restates the four per-image training losses (RetinaNet.py:357-474,
SSD300.py:345-453, YOLOv3.py:115-318, FCOS.py:153-187,266-348) for the loss-forward kernels;
its pins are the hand-derived tiny cases of tests/test_oracle_loss.py and the frozen values in
tests/golden/loss_*.npz -- equally unpinned against TensorFlow itself.
"""
