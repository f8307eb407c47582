import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import tartan_oracle as O
from tartangan_b200.trainers.cnn import CNNTrainer
from tartangan_b200.trainers.gan import make_trainer
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
torch.manual_seed(0)
t = make_trainer(CNNTrainer, config='64', batch_size=16, precision=prec)
cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
orc = O.OracleTrainer('cnn', O.SPECS['64'], cpu(t.g), cpu(t.target_g), cpu(t.d), 16)
orc2 = O.OracleTrainer('cnn', O.SPECS['64'], cpu(t.g), cpu(t.target_g), cpu(t.d), 16)
torch.set_num_threads(4)
for s in range(20):
    imgs = O.tartan_batch(1234 + s, 16, 64)
    torch.manual_seed(2000 + s); ref = orc.train_batch(imgs)
    torch.manual_seed(2000 + s); got = t.train_batch(imgs)
    print(s, ' '.join(f"{k}: {ref[k]:.4f}/{got[k]:.4f}" for k in ref))
