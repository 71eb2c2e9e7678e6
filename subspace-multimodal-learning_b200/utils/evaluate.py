"""On-device evaluation for the hot-path models.

Reference: ``utils/metrics.py:9-25`` (``compute_avg_metrics``: sklearn / imblearn on host copies) and ``:38-70`` (``epochVal``: grows
``groundTruth`` / ``activations`` with ``torch.cat`` every batch, one allocation + copy per step).  Here the epoch's labels and
softmax outputs land in buffers allocated once, every metric is computed on the device from the confusion matrix and the
rank statistic, and the host reads seven scalars at the end -- one synchronisation per evaluation, none per batch.  Same
return tuple and definitions (macro averages over the labels that occur, one-vs-rest AUC, 0 when a class is absent as the
reference's ``except ValueError`` branch does), so ``train_test.py:263-290`` can call it unchanged.
"""
import threading

import torch
import torch.nn.functional as F


def _avg_ranks(x):
    """average ranks (1-based, ties share the mean rank) of a 1-D tensor, on its device"""
    order = torch.argsort(x, stable=True)
    xs = x[order]
    n = x.numel()
    idx = torch.arange(n, device=x.device)
    new = torch.ones(n, dtype=torch.bool, device=x.device)
    new[1:] = xs[1:] != xs[:-1]
    start = torch.cummax(torch.where(new, idx, torch.zeros_like(idx)), 0).values          # first index of my tie group
    last = torch.ones(n, dtype=torch.bool, device=x.device)
    last[:-1] = new[1:]
    end = torch.flip(torch.cummin(torch.flip(torch.where(last, idx, torch.full_like(idx, n)), [0]), 0).values, [0])
    r = (start + end).double() * 0.5 + 1.0
    out = torch.empty(n, dtype=torch.float64, device=x.device)
    out[order] = r
    return out


def avg_metrics_device(ground_truth, activations):
    """(acc, f1_macro, auc_ovr, balanced_acc, sens_macro, spec_macro, prec_macro) as a 7-element float64 device tensor"""
    y = ground_truth.long().view(-1)
    act = activations.double()
    n, C = act.shape
    pred = act.argmax(-1)
    cm = torch.zeros(C, C, dtype=torch.float64, device=act.device)
    cm.view(-1).index_add_(0, y * C + pred, torch.ones(n, dtype=torch.float64, device=act.device))
    tp = cm.diag()
    support, predicted = cm.sum(1), cm.sum(0)
    present = (support + predicted) > 0                       # sklearn / imblearn: labels = union of y_true and y_pred
    in_true = support > 0
    npres = present.sum().clamp(min=1)
    recall = torch.where(support > 0, tp / support.clamp(min=1), torch.zeros_like(tp))
    prec = torch.where(predicted > 0, tp / predicted.clamp(min=1), torch.zeros_like(tp))
    f1 = torch.where(prec + recall > 0, 2 * prec * recall / (prec + recall).clamp(min=1e-300), torch.zeros_like(tp))
    tn = n - support - predicted + tp
    spec = torch.where(n - support > 0, tn / (n - support).clamp(min=1), torch.zeros_like(tp))
    acc = tp.sum() / n
    f1_macro = (f1 * present).sum() / npres
    bac = (recall * in_true).sum() / in_true.sum().clamp(min=1)
    sens_macro = (recall * present).sum() / npres
    spec_macro = (spec * present).sum() / npres
    prec_macro = (prec * present).sum() / npres
    # one-vs-rest AUC from the Mann-Whitney statistic; undefined (-> 0, as the reference's ValueError branch) unless every
    # class has a positive and a negative example
    aucs = []
    for c in range(C):
        pos = (y == c)
        npos = pos.sum().double()
        r = _avg_ranks(act[:, c])
        u = (r * pos).sum() - npos * (npos + 1) / 2
        aucs.append(u / (npos * (n - npos)).clamp(min=1))
    ok = ((support > 0) & (support < n)).all()
    auc = torch.where(ok, torch.stack(aucs).mean(), torch.zeros((), dtype=torch.float64, device=act.device))
    return torch.stack([acc, f1_macro, auc, bac, sens_macro, spec_macro, prec_macro])


def compute_avg_metrics(groundTruth, activations):
    """drop-in for utils/metrics.py:9-25 (returns python floats)"""
    return tuple(float(v) for v in avg_metrics_device(groundTruth, activations).cpu())


_LABEL_COL = {"diag2021": 5, "grade": 4, "subtype": 7}


def epochVal(model, dataLoader, args):
    """drop-in for utils/metrics.py:38-70: evaluates `model` (DeformPathomicNet interface) over `dataLoader`"""
    training = model.training
    model.eval()
    dev = next(model.parameters()).device
    gt = act = None
    fill = 0
    with torch.no_grad():
        for x_path, x_omic, x_omic_tumor, x_omic_immune, label in dataLoader:
            x_path, x_omic_tumor, x_omic_immune, label = (t.to(dev, non_blocking=True) for t in (x_path, x_omic_tumor, x_omic_immune, label))
            _, _, _, logits, _, _, _ = model(x_path=x_path, x_omic=x_omic, x_omic_tumor=x_omic_tumor, x_omic_immune=x_omic_immune)
            out = F.softmax(logits[2], dim=1)
            b = out.shape[0]
            if act is None:
                try:
                    cap = max(len(dataLoader) * b, b)
                except TypeError:
                    cap = 64 * b
                act = torch.empty(cap, out.shape[1], device=dev, dtype=out.dtype)
                gt = torch.empty(cap, device=dev, dtype=torch.int64)
            if fill + b > act.shape[0]:                       # loader without a length: amortised doubling
                act = torch.cat([act, torch.empty_like(act)]); gt = torch.cat([gt, torch.empty_like(gt)])
            act[fill:fill + b] = out
            gt[fill:fill + b] = label[:, _LABEL_COL[args.task_type]]
            fill += b
    res = compute_avg_metrics(gt[:fill], act[:fill])
    model.train(training)
    return res


class AsyncEvaluator:
    """Periodic evaluation that does not stall the train loop (reference: train_test.py:195-250 evaluates on rank 0
    every 10 iterations *inside* the loop, so every rank waits for it at the next collective).

    ``submit(model, tag)`` snapshots the weights into a shadow model on a side stream (ordered after everything already
    enqueued on the training stream) and runs ``epochVal`` over ``make_loader()`` on that stream from a worker thread;
    training continues on its own stream.  ``poll()`` returns the finished ``(tag, metrics)`` pairs, ``wait()`` blocks for the
    one in flight.  One evaluation at a time: ``submit`` while one is running returns False (the reference would have blocked).
    """

    def __init__(self, shadow_model, make_loader, args):
        self.shadow, self.make_loader, self.args = shadow_model.eval(), make_loader, args
        self.dev = next(shadow_model.parameters()).device
        self.stream = torch.cuda.Stream(device=self.dev)
        self.thread, self.done, self.error = None, [], None
        self._lock = threading.Lock()

    def busy(self):
        return self.thread is not None and self.thread.is_alive()

    def submit(self, model, tag=None):
        if self.busy():
            return False
        src = (model.module if hasattr(model, "module") else model).state_dict()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.stream), torch.no_grad():
            self.stream.wait_event(ev)
            for k, v in self.shadow.state_dict().items():
                v.copy_(src[k], non_blocking=True)
        # the optimizer must not overwrite the weights while the snapshot copies read them
        torch.cuda.current_stream(self.dev).wait_stream(self.stream)

        def run():
            try:
                torch.cuda.set_device(self.dev)
                with torch.cuda.stream(self.stream):
                    res = epochVal(self.shadow, self.make_loader(), self.args)
                with self._lock:
                    self.done.append((tag, res))
            except Exception as e:                           # noqa: BLE001 - surfaced by poll() / wait()
                self.error = e
        self.thread = threading.Thread(target=run, daemon=True)
        self.thread.start()
        return True

    def poll(self):
        if self.error is not None:
            e, self.error = self.error, None
            raise e
        with self._lock:
            out, self.done = self.done, []
        return out

    def wait(self):
        if self.thread is not None:
            self.thread.join()
        return self.poll()
