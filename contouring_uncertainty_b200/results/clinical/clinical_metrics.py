"""Mirror of contour_uncertainty/results/clinical/clinical_metrics.py::ClinicalMetrics (:31-143), numbers only: the data frames
of the instant / view / patient metric processors and the dataset-level figures the reference logs (reject rates, Pearson
correlation, MAE, UCE, adaptive UCE).  Plotting, CSV dumps next to the working directory and the Comet logger stay with the
reference."""
from collections import defaultdict
from pathlib import Path
from typing import List, Optional

import numpy as np
import pandas as pd

from ...data.config import BatchResult
from ..utils.calibration import metric_calibration, metric_correlation
from .instant import AreaError, InstantMetric
from .patient import PatientMetric, Volume
from .view import FAC, GLS, ViewMetric


class ClinicalMetrics:
    def __init__(self, instant_metrics: List[InstantMetric] = (AreaError(),), view_metrics: List[ViewMetric] = (FAC(), GLS(),),
                 patient_metrics: List[PatientMetric] = (Volume(),), output_dir: Optional[Path] = None):
        self.instant_metrics = instant_metrics or []
        self.view_metrics = view_metrics or []
        self.patient_metrics = patient_metrics or []
        self.metric_results = {}
        self.output_dir = Path(output_dir) if output_dir is not None else None
        if self.output_dir is not None:
            self.output_dir.mkdir(parents=True, exist_ok=True)
        self.instant_df = self.view_df = self.patient_df = self.volume_df = None

    def correlation(self, df: pd.DataFrame, metric_name: str, x_str: str = 'gt', y_str: str = 'pred'):
        """plot_correlation, the numbers (:156-190)."""
        r = metric_correlation(np.array(df[f'{metric_name}_{x_str}'], dtype=float), np.array(df[f'{metric_name}_{y_str}'], dtype=float))
        self.metric_results[f'{metric_name}_{x_str}_{y_str}_correlation'] = r["correlation"]
        self.metric_results[f'{metric_name}_{x_str}_{y_str}_mae'] = r["mae"]

    def calibration(self, df: pd.DataFrame, metric_name: str):
        """plot_calibration, the numbers (:204-226 and the logging below it)."""
        reject = np.array(df[f'{metric_name}_reject'], dtype=bool) if f'{metric_name}_reject' in df.columns else None
        r = metric_calibration(np.array(df[f'{metric_name}_std'], dtype=float), np.array(df[f'{metric_name}_error'], dtype=float), reject)
        self.metric_results[f'{metric_name}_uce'] = r["uce"]
        self.metric_results[f'{metric_name}_a_uce'] = r["a_uce"]

    def on_predict_epoch_end(self, trainer=None, pl_module=None, outputs: List[List[BatchResult]] = None) -> dict:
        """clinical_metrics.py:51-143.  ``outputs[0]``: the BatchResults of all views (``id`` = 'patient/view')."""
        self.metric_results = {}
        result_dict = defaultdict(dict)
        for res in outputs[0]:
            patient_id, view_id = res.id.split('/')
            result_dict[patient_id][view_id] = res
        if self.instant_metrics:
            self.instant_df = pd.concat([metric(outputs[0]) for metric in self.instant_metrics])
            self.correlation(self.instant_df, 'Area')
            self.correlation(self.instant_df, 'Area', x_str='pred', y_str='mean')
            self.calibration(self.instant_df, 'Area')
        if self.view_metrics:
            self.view_df = pd.concat([metric(outputs[0]) for metric in self.view_metrics], axis=1)
            for metric in ('FAC', 'GLS'):
                if f'{metric}_reject' in self.view_df.columns:
                    reject = np.array(self.view_df[f'{metric}_reject'], dtype=bool)
                    self.metric_results[f'{metric}_reject'] = reject.sum() / np.size(reject)
                    self.correlation(self.view_df, metric)
                    self.correlation(self.view_df, metric, x_str='pred', y_str='mean')
                    self.calibration(self.view_df, metric)
        if self.patient_metrics:
            self.patient_df = pd.concat([metric(result_dict) for metric in self.patient_metrics])
            if len(self.patient_df):
                for metric in ('EF', 'ESV', 'EDV'):
                    self.correlation(self.patient_df, metric)
                    self.correlation(self.patient_df, metric, x_str='pred', y_str='mean')
                    self.calibration(self.patient_df, metric)
                reject = np.array(self.patient_df['EF_reject'], dtype=bool)
                self.metric_results['EF_reject'] = reject.sum() / np.size(reject)
                self.volume_df = self.merge_volume_df(self.patient_df)
                self.correlation(self.volume_df, 'Volume')
                self.correlation(self.volume_df, 'Volume', x_str='pred', y_str='mean')
                self.calibration(self.volume_df, 'Volume')
        if self.output_dir is not None:
            for name in ("instant_df", "view_df", "patient_df", "volume_df"):
                df = getattr(self, name)
                if df is not None:
                    df.to_csv(self.output_dir / f"{name}.csv")
            pd.DataFrame([self.metric_results]).to_csv(self.output_dir / 'clinical_metrics.csv')
        if trainer is not None and getattr(trainer, "logger", None) is not None:
            trainer.logger.log_metrics(self.metric_results)
        return self.metric_results

    @staticmethod
    def merge_volume_df(patient_df: pd.DataFrame):
        """clinical_metrics.py:145-154."""
        esv = patient_df.filter(regex='ESV')
        esv.index = [f'{patient}/ES' for patient in esv.index]
        esv.columns = [f'Volume_{col.replace("ESV_", "")}' for col in esv.columns]
        edv = patient_df.filter(regex='EDV')
        edv.index = [f'{patient}/ED' for patient in edv.index]
        edv.columns = [f'Volume_{col.replace("EDV_", "")}' for col in edv.columns]
        return pd.concat([esv, edv])
