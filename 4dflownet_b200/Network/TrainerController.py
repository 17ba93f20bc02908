"""Drop-in mirror of the reference's Network/TrainerController.py.

Same constructor signature (TrainerController.py:18), same public methods
(`init_model_dir` :158, `train_step` :210, `test_step` :228, `train_network` :263,
`save_best_model` :347, `restore_model` :365, `quicksave` :415) and the same nine running
metrics (:52-63).  What changed is underneath: the Keras graph, `tf.GradientTape` and the
Keras Adam are one libsr4d engine (forward + loss/metric + backward kernels, one fused Adam
kernel over the flat parameter buffer), and when `torch.distributed` is initialised the
batch shard's summed gradients are all-reduced (SUM) once per step on the flat buffer
before Adam — the data-parallel scheme of the north star.

Step semantics reproduced exactly (TrainerController.py:209-257):
  loss vector (B,) = fluid + non-fluid masked SSE  (+ l2 scalar broadcast onto each entry)
  gradients       = d/dw  sum_b loss_b            (sum, not mean; hence B_global * d(l2)/dw)
  metrics         = running means over every sample seen in the epoch.
"""
import datetime
import os
import pickle
import shutil
import time

import numpy as np
import torch

from .SR4DFlowNet import SR4DFlowModel
from . import utility
from .. import parallel

L2_COEFF = 5e-7    # SR4DFlowNet.py:99 kernel_regularizer=l2(5e-7)


class Mean:
    """tf.keras.metrics.Mean: running mean over all values passed to update_state."""

    def __init__(self, name=None, before_access=None):
        self.name = name
        self.total, self.count = 0.0, 0
        # hook run before every access: the controller folds the metrics of a step whose read-back is still in flight
        self._before = before_access

    def update_state(self, values):
        if self._before:
            self._before()
        v = np.asarray(values, dtype=np.float64).reshape(-1)
        self.total += float(v.sum())
        self.count += v.size

    def result(self):
        if self._before:
            self._before()
        return self.total / self.count if self.count else 0.0

    def reset_states(self):
        if self._before:
            self._before()
        self.total, self.count = 0.0, 0


class _LearningRate:
    def __init__(self, v):
        self.v = float(v)

    def numpy(self):
        return self.v

    def assign(self, v):
        self.v = float(v)

    def __float__(self):
        return self.v


class AdamOptimizer:
    """The slice of tf.keras.optimizers.Adam the reference touches (TrainerController.py:73,
    225,269,359,391): `lr`, `iterations`, `weights` = [iterations, m_0..m_n, v_0..v_n],
    `set_weights`, and an apply step.  State lives in the engine's flat m / v buffers.

    ORDER.  The m / v lists follow this package's parameter table, i.e. Keras layer-CREATION order (conv3d,
    conv3d_1, ...).  A Keras functional model lists `trainable_variables` / `optimizer.weights` by graph depth
    instead (parallel branches interleave), which cannot be reproduced here without TensorFlow: an `optimizer.pkl`
    written by the reference's own training run is therefore NOT restorable through `set_weights` (named `.h5` weight
    files are: they are matched by layer name).  Files this package wrote round-trip.  `set_weights` validates every
    shape before it touches the state, so a foreign list fails cleanly instead of half-overwriting the moments."""

    def __init__(self, engine, lr=1e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.engine = engine
        self.lr = _LearningRate(lr)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.iterations = 0

    @property
    def learning_rate(self):
        return self.lr

    @property
    def weights(self):
        m = [v.detach().cpu().numpy().copy() for _, v in self.engine.tensor_views(self.engine.adam_m)]
        v = [v.detach().cpu().numpy().copy() for _, v in self.engine.tensor_views(self.engine.adam_v)]
        return [np.int64(self.iterations)] + m + v

    def get_weights(self):
        return self.weights

    def set_weights(self, weights):
        n = len(self.engine.table)
        if len(weights) != 1 + 2 * n:
            raise ValueError(f"expected {1 + 2 * n} optimizer tensors, got {len(weights)}")
        mviews = self.engine.tensor_views(self.engine.adam_m)
        vviews = self.engine.tensor_views(self.engine.adam_v)
        staged = [torch.as_tensor(np.asarray(w, dtype=np.float32)) for w in weights[1:]]
        for (name, view), w in zip(mviews + vviews, staged):           # validate all shapes before the first copy
            if tuple(w.shape) != tuple(view.shape):
                raise ValueError(f"optimizer slot of {name}: shape {tuple(w.shape)} != {tuple(view.shape)}; the list must be "
                                 "[iterations, m..., v...] in this package's table order (layer-creation order), not "
                                 "the graph-depth order of a Keras-written optimizer.pkl")
        self.iterations = int(weights[0])
        for (_, view), w in zip(mviews + vviews, staged):
            view.copy_(w)

    def apply(self, global_batch):
        """apply_gradients on the engine's gradient buffer (TrainerController.py:225)."""
        self.iterations += 1
        self.engine.adam_step(float(self.lr), self.iterations, global_batch * 2.0 * L2_COEFF,
                              self.beta_1, self.beta_2, self.epsilon)

    def apply_counted(self, tail_index):
        """Same, with the global batch size read on the device from the metric tail of the gradient buffer (what the
        all-reduce left there): no host synchronisation between the collective and the update."""
        self.iterations += 1
        self.engine.adam_step_counted(float(self.lr), self.iterations, 2.0 * L2_COEFF, tail_index,
                                      self.beta_1, self.beta_2, self.epsilon)


def _squeeze_last(a):
    return a[..., 0] if a.shape[-1] == 1 and a.ndim == 5 else a


class TrainerController:
    def __init__(self, patch_size, res_increase, initial_learning_rate=1e-4, quicksave_enable=True,
                 network_name='4DFlowNet', low_resblock=8, hi_resblock=4, max_batch=8, device=None, seed=None):
        self.div_weight = 0          # divergence loss is disabled in the reference (:23,:121)
        self.non_fluid_weight = 1
        self.res_increase = res_increase
        self.patch_size = patch_size
        self.QUICKSAVE_ENABLED = quicksave_enable
        self.network_name = network_name

        self.model = SR4DFlowModel(patch_size, res_increase, low_resblock, hi_resblock, max_batch=max_batch,
                                   training=True, device=device, seed=seed)
        self.engine = self.model.engine

        self._pending = None      # token of a train step whose metric read-back has been enqueued but not folded yet
        self.loss_metrics = dict((k, Mean(name=k, before_access=self._fold_pending)) for k in (
            'train_loss', 'val_loss', 'train_accuracy', 'val_accuracy', 'train_mse', 'val_mse',
            'train_div', 'val_div', 'l2_reg_loss'))
        self.accuracy_metric = 'val_loss'
        print(f"Divergence loss2 * {self.div_weight}")
        print(f"Accuracy metric: {self.accuracy_metric}")

        self.learning_rate = initial_learning_rate
        self.optimizer = AdamOptimizer(self.engine, lr=self.learning_rate)
        self._last = None    # device tensors of the most recent step (per_sample, l2)
        self._metric_tail = None

    # ---- loss / metric entry points with the reference's names -------------------------------
    def loss_function(self, y_true, y_pred, mask):
        """(total_loss[B], mse[B], 0) — TrainerController.py:84-127."""
        per = self.engine.loss_metrics(y_pred, y_true[..., 0], y_true[..., 1], y_true[..., 2], mask)
        per = per.cpu().numpy()
        return per[:, 0], per[:, 1], 0

    def accuracy_function(self, y_true, y_pred, mask):
        """relative speed error in % per sample — TrainerController.py:143-150."""
        per = self.engine.loss_metrics(y_pred, y_true[..., 0], y_true[..., 1], y_true[..., 2], mask)
        return per[:, 2].cpu().numpy()

    def calculate_mse(self, u, v, w, u_pred, v_pred, w_pred):
        """Voxel-wise squared error summed over the three components (TrainerController.py:152-156).  Convenience
        mirror only: the training path computes it inside the fused loss kernel."""
        t = [torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(self.engine.device, torch.float32)
             for a in (u, v, w, u_pred, v_pred, w_pred)]
        return (t[3] - t[0]) ** 2 + (t[4] - t[1]) ** 2 + (t[5] - t[2]) ** 2

    def calculate_and_update_metrics(self, hires, predictions, mask, metric_set):
        """TrainerController.py:241-257: loss / mse / accuracy of a batch folded into the running means; returns the
        (B,) loss vector (with the scalar l2 added for the train set)."""
        hires = torch.as_tensor(np.asarray(hires) if not torch.is_tensor(hires) else hires)
        per = self.engine.loss_metrics(predictions, hires[..., 0], hires[..., 1], hires[..., 2], mask)
        l2 = self.calculate_regularizer_loss() if metric_set == 'train' else None
        return self._update_metrics(per.cpu().numpy(), l2, metric_set)

    def calculate_regularizer_loss(self):
        """5e-7 * sum over the conv kernels of sum(w^2) — TrainerController.py:129-141."""
        tot = 0.0
        for (name, view) in self.engine.tensor_views():
            if name.endswith("kernel"):
                tot += float((view.double() ** 2).sum())
        return L2_COEFF * tot

    # ---- steps --------------------------------------------------------------------------------
    def _tail(self):
        if self._metric_tail is not None and self._metric_tail.world != parallel.world_size():
            self._fold_pending()             # a read-back of the old tail object is still in flight
        if self._metric_tail is None or self._metric_tail.world != parallel.world_size():
            self._metric_tail = parallel.MetricTail(self.engine.grad_tail, self.engine.max_batch)
        return self._metric_tail

    def train_step_async(self, data_pairs):
        """Enqueue forward + loss + backward + the step's ONE collective + Adam; returns device views
        (per_sample (B,4), l2 (1,)) of this rank's metrics without synchronising.

        Data parallel: every rank holds a shard of the global batch and leaves the SUM of its samples'
        gradients in the flat buffer; the per-sample metrics, the l2 value and the local batch size go into the
        rank's slot of the metric tail behind the gradients, so ONE all-reduce(SUM) of `grads_full` delivers the
        global gradient, gathers the metrics of every sample and the global batch size; Adam then runs identically on
        every rank with the L2 gradient of the GLOBAL batch (TrainerController.py:223,249) read from the tail on the
        device."""
        u, v, w, u_mag, v_mag, w_mag, u_hr, v_hr, w_hr, venc, mask = data_pairs
        tail = self._tail()
        per, l2 = tail.begin(len(u))
        self.engine.train_fwd_bwd([u, v, w, u_mag, v_mag, w_mag], [_squeeze_last(a) for a in (u_hr, v_hr, w_hr)], mask,
                                  per_out=per, l2_out=l2)
        parallel.allreduce_gradients(self.engine.grads_full)
        self.optimizer.apply_counted(tail.count_index)
        self._last = (per, l2)
        return per, l2

    def train_step(self, data_pairs):
        """TrainerController.py:209-225.  The step's metrics come back through one small device->host copy that is only
        waited for when somebody looks at the running means (or at the next train_step, after that step has been
        enqueued): the GPU never idles while the host folds numbers."""
        self.train_step_async(data_pairs)
        token = self._tail().read_begin()
        self._fold_pending()                 # the previous step's copy finished long ago
        self._pending = token

    def _fold_pending(self):
        if self._pending is None:
            return
        token, self._pending = self._pending, None
        per, l2, _ = self._tail().read_end(token)     # running means cover every sample of the global batch
        self._update_metrics(per, l2, 'train')

    def test_step(self, data_pairs):
        u, v, w, u_mag, v_mag, w_mag, u_hr, v_hr, w_hr, venc, mask = data_pairs
        predictions = self.model([u, v, w, u_mag, v_mag, w_mag], training=False)
        tail = self._tail()
        per, _ = tail.begin(len(u))
        self.engine.loss_metrics(predictions, *[_squeeze_last(a) for a in (u_hr, v_hr, w_hr)], mask, per_out=per)
        tail.exchange()
        self._update_metrics(tail.read()[0], None, 'val')
        return predictions

    def _update_metrics(self, per, l2, metric_set):
        """calculate_and_update_metrics (:241-257): l2 is added to every entry of the loss
        vector for the train set only."""
        loss = per[:, 0].astype(np.float64)
        if metric_set == 'train':
            self.loss_metrics['l2_reg_loss'].update_state(l2)
            loss = loss + l2
        self.loss_metrics[f'{metric_set}_loss'].update_state(loss)
        self.loss_metrics[f'{metric_set}_mse'].update_state(per[:, 1])
        self.loss_metrics[f'{metric_set}_div'].update_state(0.0)
        self.loss_metrics[f'{metric_set}_accuracy'].update_state(per[:, 2])
        return loss

    def reset_metrics(self):
        for m in self.loss_metrics.values():
            m.reset_states()

    # ---- model directory, logging ---------------------------------------------------------------
    def init_model_dir(self, root="../models"):
        timestamp = datetime.datetime.now().strftime("%Y%m%d-%H%M")
        if parallel.world_size() > 1:
            # one directory per run: every rank uses rank 0's timestamp; only rank 0 writes into it
            import torch.distributed as dist
            box = [timestamp]
            dist.broadcast_object_list(box, src=0)
            timestamp = box[0]
        self.unique_model_name = f'{self.network_name}_{timestamp}'
        self.model_dir = f"{root}/{self.unique_model_name}"
        self.model_path = f"{self.model_dir}/{self.network_name}"
        self.train_writer = self.val_writer = None
        self.logfile = self.model_dir + '/loss.csv'
        if parallel.is_main():
            os.makedirs(self.model_dir, exist_ok=True)
            self._prepare_logfile_and_summary()
        parallel.barrier()

    def _log(self, text):
        if parallel.is_main():
            utility.log_to_file(self.logfile, text)

    def sync_ranks(self):
        """Data parallel: rank 0's weights, Adam moments and iteration count on every rank (after the random
        initialisation and after restore_model); the derived tensor-core weight images are rebuilt."""
        if parallel.world_size() == 1:
            return
        for t in (self.engine.params, self.engine.adam_m, self.engine.adam_v):
            parallel.broadcast_(t, 0)
        it = torch.tensor([self.optimizer.iterations, 0], dtype=torch.int64, device=self.engine.device)
        it[1] = int(np.float32(float(self.optimizer.lr)).view(np.int32))
        parallel.broadcast_(it, 0)
        self.optimizer.iterations = int(it[0].item())
        self.optimizer.lr.assign(float(np.int32(int(it[1].item())).view(np.float32)))
        self.engine.params_changed()

    def _prepare_logfile_and_summary(self):
        self.train_writer = self.val_writer = None
        try:    # TensorBoard scalars when the writer is available (tensorboard package)
            from torch.utils.tensorboard import SummaryWriter
            self.train_writer = SummaryWriter(self.model_dir + '/tensorboard/train')
            self.val_writer = SummaryWriter(self.model_dir + '/tensorboard/validate')
        except Exception:
            pass
        self.logfile = self.model_dir + '/loss.csv'
        utility.log_to_file(self.logfile, f'Network: {self.network_name}\n')
        utility.log_to_file(self.logfile, f'Initial learning rate: {self.learning_rate}\n')
        utility.log_to_file(self.logfile, f'Accuracy metric: {self.accuracy_metric}\n')
        utility.log_to_file(self.logfile, f'Divergence weight: {self.div_weight}\n')
        stat_names = ','.join(self.loss_metrics.keys())
        utility.log_to_file(self.logfile, f'epoch, {stat_names}, learning rate, elapsed (sec), best_model, '
                                          'benchmark_err, benchmark_rel_err, benchmark_mse, benchmark_divloss\n')

    def _update_summary_logging(self, epoch):
        if self.train_writer is None:
            return
        self.train_writer.add_scalar(f"{self.network_name}/learning_rate", float(self.optimizer.lr), epoch)
        for k, m in self.loss_metrics.items():
            if k.startswith('train_'):
                self.train_writer.add_scalar(f"{self.network_name}/{k[6:]}", m.result(), epoch)
            elif k.startswith('val_'):
                self.val_writer.add_scalar(f"{self.network_name}/{k[4:]}", m.result(), epoch)

    # ---- main loop (TrainerController.py:263-343) -------------------------------------------------
    def train_network(self, trainset, valset, n_epoch, testset=None):
        print("==================== TRAINING =================")
        print(f'Learning rate {self.optimizer.lr.numpy():.7f}')
        print(f"Start training at {time.ctime()} - {self.unique_model_name}\n")
        start_time = time.time()
        previous_loss = np.inf
        total_batch_train = len(trainset) if hasattr(trainset, "__len__") else -1
        total_batch_val = len(valset) if hasattr(valset, "__len__") else -1
        for epoch in range(n_epoch):
            self.reset_metrics()
            start_loop = time.time()
            for i, data_pairs in enumerate(trainset):
                self.train_step(data_pairs)
                print(f"\rEpoch {epoch+1} Train batch {i+1}/{total_batch_train} | loss: "
                      f"{self.loss_metrics['train_loss'].result():.5f} ({self.loss_metrics['train_accuracy'].result():.1f} %)"
                      f" - {time.time()-start_loop:.1f} secs", end='')
            for i, data_pairs in enumerate(valset):
                self.test_step(data_pairs)
                print(f"\rEpoch {epoch+1} Validation batch {i+1}/{total_batch_val} | loss: "
                      f"{self.loss_metrics['val_loss'].result():.5f} ({self.loss_metrics['val_accuracy'].result():.1f} %)"
                      f" - {time.time()-start_loop:.1f} secs", end='')
            message = (f"\rEpoch {epoch+1} Train loss: {self.loss_metrics['train_loss'].result():.5f} "
                       f"({self.loss_metrics['train_accuracy'].result():.1f} %), Val loss: "
                       f"{self.loss_metrics['val_loss'].result():.5f} ({self.loss_metrics['val_accuracy'].result():.1f} %)"
                       f" - {time.time()-start_loop:.1f} secs")
            loss_str = ','.join(f'{m.result():.5f}' for m in self.loss_metrics.values())
            log_line = f"{epoch+1},{loss_str},{self.optimizer.lr.numpy():.6f},{time.time()-start_loop:.1f}"
            self._update_summary_logging(epoch)
            if self.loss_metrics[self.accuracy_metric].result() < previous_loss:
                self.save_best_model()
                previous_loss = self.loss_metrics[self.accuracy_metric].result()
                message += ' **'
                log_line += ',**'
                if self.QUICKSAVE_ENABLED and testset is not None:
                    q = [float(np.mean(x)) for x in self.quicksave(testset, epoch + 1)]
                    message += f' Benchmark loss: {q[0]:.5f} ({q[1]:.1f} %)'
                    log_line += f', {q[0]:.7f}, {q[1]:.2f}%, {q[2]:.7f}, {q[3]:.7f}'
            print(message)
            self._log(log_line + "\n")
        hrs, mins, secs = utility.calculate_time_elapsed(start_time)
        message = (f"\nTraining {self.network_name} completed! - name: {self.unique_model_name}"
                   f"\nTotal training time: {hrs} hrs {mins} mins {secs} secs."
                   f"\nFinished at {time.ctime()}\n==================== END TRAINING =================")
        self._log(message)
        print(message)

    # ---- checkpoints (TrainerController.py:347-394) -------------------------------------------------
    def save_latest_model(self, epoch):
        if epoch > 0 and epoch % 10 == 0:
            if parallel.is_main():
                self.model.save(f'{self.model_path}-latest.h5')
                print(f'Saving current model - {time.ctime()}\n')
            parallel.barrier()

    def save_best_model(self):
        """Rank 0 writes (weights are identical on every rank after the all-reduced Adam step); the others wait."""
        if parallel.is_main():
            self.model.save(f'{self.model_path}-best.h5')
            with open(f'{self.model_dir}/optimizer.pkl', 'wb') as f:
                pickle.dump(self.optimizer.weights, f)
        parallel.barrier()

    def restore_model(self, old_model_dir, old_model_file):
        with open(f"{old_model_dir}/optimizer.pkl", 'rb') as f:
            opt_weights = pickle.load(f)
        self.optimizer.set_weights(opt_weights)
        self.model.load_weights(f"{old_model_dir}/{old_model_file}")

    def quicksave(self, testset, epoch_nr):
        """First batch of the benchmark set -> quicksave_<network_name>.h5 (:415-454).  The benchmark iterator yields
        GLOBAL batches (it is not sharded); rank 0 alone predicts it, in chunks of at most the engine's max_batch (under
        data parallelism the engine is sized for batch_size / ranks), and writes the file; the other ranks wait."""
        from . import h5util
        if not parallel.is_main():
            parallel.barrier()
            z = np.zeros(1, dtype=np.float32)
            return z, z, z, z
        for data_pairs in testset:
            u, v, w, u_mag, v_mag, w_mag, u_hr, v_hr, w_hr, venc, mask = data_pairs
            n, mb = len(u), self.engine.max_batch
            preds_l, per_l = [], []
            for lo in range(0, n, mb):
                sl = slice(lo, min(lo + mb, n))
                pd = self.model([a[sl] for a in (u, v, w, u_mag, v_mag, w_mag)])
                per_l.append(self.engine.loss_metrics(pd, *[_squeeze_last(a[sl]) for a in (u_hr, v_hr, w_hr)], mask[sl]).cpu().numpy())
                preds_l.append(pd.cpu().numpy())
            per = np.concatenate(per_l, axis=0)
            preds = np.concatenate(preds_l, axis=0)
            break
        fn = f"quicksave_{self.network_name}.h5"
        h5util.save_predictions(self.model_dir, fn, "epoch", np.asarray([epoch_nr]), compression='gzip')
        p = np.expand_dims(preds, 0)
        for i, c in enumerate("uvw"):
            h5util.save_predictions(self.model_dir, fn, c, p[..., i], compression='gzip')
        if epoch_nr == 1:
            for name, a in (("lr_u", u), ("lr_v", v), ("lr_w", w)):
                h5util.save_predictions(self.model_dir, fn, name, np.asarray(a), compression='gzip')
            for name, a in (("hr_u", u_hr), ("hr_v", v_hr), ("hr_w", w_hr)):
                h5util.save_predictions(self.model_dir, fn, name, np.squeeze(np.asarray(a), -1), compression='gzip')
            h5util.save_predictions(self.model_dir, fn, "venc", np.asarray(venc), compression='gzip')
            h5util.save_predictions(self.model_dir, fn, "mask", np.asarray(mask), compression='gzip')
        parallel.barrier()
        return per[:, 0], per[:, 2], per[:, 1], np.zeros_like(per[:, 0])
