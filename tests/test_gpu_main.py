"""YAML-driven entry points (enflow/main.py:280-288) on the CUDA path: train, checkpoint, resume, generate."""
import os

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(tmp_path, name, **over):
    cfg = yaml.safe_load(open(os.path.join(ROOT, 'example', name)))
    cfg['dynamics']['checkpoint_path'] = str(tmp_path / 'model.cpt')
    for k, v in over.items():
        sec, key = k.split('__')
        cfg[sec][key] = v
    path = tmp_path / name
    yaml.safe_dump(cfg, open(path, 'w'))
    return str(path)


def test_train_checkpoint_resume_generate(tmp_path):
    from enflow_b200.main import Main
    os.chdir(tmp_path)
    train = _cfg(tmp_path, 'train_synthetic.yaml', dataset__num_mols=128, training__num_epochs=2)
    torch.manual_seed(3)                           # initial weights, shuffling and dequantisation noise
    loss1 = Main()(train)
    ck = torch.load(tmp_path / 'model.cpt', weights_only=False)
    # checkpoint schema of enflow/main.py:236-250
    assert set(ck) >= {'epoch', 'model_state_dict', 'optimizer_state_dict', 'node_nf', 'hidden_nf', 'softening',
                       'lj_kBT', 'integrator', 'n_iter', 'dt'}
    assert ck['epoch'] == 1 and ck['n_iter'] == 5 and ck['integrator'] == 'lf'
    assert list(ck['model_state_dict'])[0] == 'networks.0.edge_nn.0.weight'
    m = Main()
    torch.manual_seed(4)
    m.setup(train)                                 # resume: hyper-parameters come from the checkpoint (main.py:100-109)
    assert m.start_epoch == 2
    loss2 = m.train()
    # fresh dequantisation noise per step makes the epoch loss of this 2-batch run noisy (the CUDA-core and tensor-core
    # modes follow the same noisy curve): only require it to stay sane after resume
    assert loss2 == loss2 and loss2 < 2.0 * loss1, (loss1, loss2)
    gen = _cfg(tmp_path, 'generate_synthetic.yaml')
    out, ok = Main()(gen)
    assert ok, 'forward(reverse(x)) must reproduce x (main.py:275-278)'
    assert os.path.exists(tmp_path / 'h.out') and os.path.exists(tmp_path / 'test_out.xyz')
    assert out.neg_ldj_mol.shape[0] == 64
    # generate from the reference's own prior: soft-LJ Langevin frames sampled on the GPU
    gen_lj = _cfg(tmp_path, 'generate_lj.yaml', dataset__n_iter=600, dataset__n_atoms=64, dataset__box=[16.0, 16.0, 16.0])
    out, ok = Main()(gen_lj)
    assert ok and out.pos.shape == (64, 3) and torch.isfinite(out.pos).all()
    assert os.path.exists(tmp_path / 'lj_log.txt') and os.path.exists(tmp_path / 'lj_traj.xyz')


def test_train_with_cuda_graph_flag(tmp_path):
    """`training.cuda_graph: true` replays the step as a CUDA graph; same data, same seed -> same first-epoch loss
    as eager launches up to the dequantisation noise stream."""
    from enflow_b200.main import Main
    os.chdir(tmp_path)
    cfg = _cfg(tmp_path, 'train_synthetic.yaml', dataset__num_mols=160, training__num_epochs=2, training__cuda_graph=True)
    torch.manual_seed(11)                      # same initial weights and shuffling for both runs
    m = Main()
    loss = m(cfg)                              # 160 molecules / 64: two graphed batches + one shorter eager batch
    assert loss == loss and loss < 1e6
    assert m._gstep is not None and not m._gstep.overflowed()
    eager = _cfg(tmp_path, 'train_synthetic.yaml', dataset__num_mols=160, training__num_epochs=2)
    (tmp_path / 'model.cpt').unlink()
    torch.manual_seed(11)
    loss_e = Main()(eager)
    assert abs(loss - loss_e) < 0.2 * abs(loss_e), (loss, loss_e)
