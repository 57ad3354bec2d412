"""The reference's workflow end to end on the GPU: generate the dataset folders, then train_model() with the flags."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('model,hidden', [('scone', [(3, 16)] * 3), ('ebli', [(3, 16)] * 3), ('bunch', [(7, 8)] * 2)])
def test_train_model_flow(tmp_path, monkeypatch, model, hidden, capsys):
    from scone_gcn_b200 import synthetic_data_gen as sdg
    from scone_gcn_b200 import trajectory_experiments as te
    monkeypatch.chdir(tmp_path)
    sdg.generate_dataset(120, 60, 'cli')
    hp = te.hyperparams(['prog', '-model', model, '-epochs', '2', '-batch_size', '16', '-data_folder_suffix', 'cli',
                         '-hidden_layers', '_'.join('%d_%d' % h for h in hidden), '-reverse', '1'])
    assert hp['epochs'] == 2.0 and hp['hidden_layers'] == hidden          # Q6: numeric flags arrive as floats
    monkeypatch.setattr(te, 'HYPERPARAMS', hp)
    np.random.seed(1030)
    net, (train_loss, train_acc, test_loss, test_acc) = te.train_model()
    out = capsys.readouterr().out
    assert 'Epoch 1 -- train loss' in out and 'standard test set:' in out and 'Reverse experiment:' in out
    assert np.isfinite([train_loss, test_loss]).all() and 0 <= train_acc <= 1 and 0 <= test_acc <= 1
    w = np.load('models/model.npy', allow_pickle=True)
    assert len(w) == len(net.weights)
    # -load_model 1 round trip (trajectory_experiments.py:464-476)
    hp2 = dict(hp, load_model=1.0, epochs=0.0)
    monkeypatch.setattr(te, 'HYPERPARAMS', hp2)
    net2, res2 = te.train_model()
    assert res2[0] == pytest.approx(train_loss, rel=1e-5) and res2[1] == pytest.approx(train_acc)
