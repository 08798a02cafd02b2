"""Two train steps of the headline model (hybrid end2end, 512x512x48, --batch slabs, `mixed`) for profiler runs:
the second step is the one a launch list / ncu capture looks at (python scripts/one_step.py [batch] [cols] [size])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import h_denseunet_b200 as hdn  # noqa: E402
from h_denseunet_b200.synthetic import synthetic_slab  # noqa: E402


class A(object):
    pass


def main():
    a = A()
    a.b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    a.input_cols = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    a.input_size = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    m = hdn.dense_rnn_net(a, precision=os.environ.get("HDN_PRECISION", "mixed"))
    m.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy])
    vol, lab = synthetic_slab(a.b, a.input_size, a.input_cols)
    net = m._net(True)
    m.train_on_batch(vol, lab)
    torch.cuda.synchronize()
    l0 = net.be.launches
    print("STEP2 begins after %d engine launches" % l0, flush=True)
    loss = m.train_step_device(net)
    torch.cuda.synchronize()
    print("step 2: loss %.5f, %d engine launches" % (loss, net.be.launches - l0))


if __name__ == "__main__":
    main()
