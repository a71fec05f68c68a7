"""InfoCollectorCallback with the surface of the reference's utils/info_collector_callback.py:5-83: it gathers the
`info['result']` of finished episodes ('Goal' / 'Out' / 'Timeout') and reports their shares per 100 episodes.

Two ways in:
  * SB3 calls `_on_step()` with `self.locals['infos']` (a list of dicts), as on the reference;
  * `collect(env)` reads the result codes of a Soccer2DVecEnv straight from its device tensors after a step - no
    Python dict per env.
With stable-baselines3 installed the class is a real BaseCallback; without it (this image) a minimal stand-in base
keeps the same methods.  The plot needs matplotlib; without it `plot_print_results` writes the same series as CSV."""
import logging

try:
    from stable_baselines3.common.callbacks import BaseCallback  # type: ignore
except Exception:  # noqa: BLE001 - SB3 is not part of this image
    class BaseCallback:  # the three members the class below relies on
        def __init__(self, verbose: int = 0):
            self.verbose = verbose
            self.locals = {}
            self.globals = {}

        def on_step(self) -> bool:
            return self._on_step()

RESULT_TYPES = ("Goal", "Out", "Timeout")


class InfoCollectorCallback(BaseCallback):
    def __init__(self):
        super().__init__()
        self.infos = []    # the info dicts of finished episodes, in order
        self.results = {}  # type -> list of percentages, one entry per 100 episodes

    def _on_step(self) -> bool:
        for info in self.locals.get("infos") or []:
            if info.get("result"):
                self.infos.append(info)
        return True

    def collect(self, env) -> int:
        """vectorised: append the results of the episodes that ended in the env's last step; returns how many"""
        from soccer2d_b200._abi import RESULT_NAMES
        codes = env.result[env.done].tolist()
        self.infos.extend({"result": RESULT_NAMES[c]} for c in codes if c)
        return len(codes)

    def reset(self):
        self.infos = []
        self.results = {}

    def update_results_dict(self, logger: logging.Logger) -> dict:
        outcomes = [info["result"] for info in self.infos]
        self.results = {kind: [] for kind in RESULT_TYPES}
        for lo in range(0, len(outcomes), 100):
            chunk = outcomes[lo:lo + 100]
            for kind in RESULT_TYPES:
                self.results[kind].append(chunk.count(kind) / len(chunk) * 100)
        logger.info(f"Results dictionary: {self.results}")
        return self.results

    def plot_print_results(self, logger: logging.Logger, file_name: str = None) -> tuple:
        self.update_results_dict(logger)
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            fig, ax = plt.subplots()
            for kind, series in self.results.items():
                ax.plot(series, label=kind)
            ax.legend()
            ax.set_xlabel("Episodes (x100)")
            ax.set_ylabel("Percentage")
            ax.set_title("Results")
            if file_name:
                fig.savefig(file_name + ".png")
            plt.close(fig)
        except ImportError:
            if file_name:
                with open(file_name + ".csv", "w") as f:
                    f.write("block," + ",".join(RESULT_TYPES) + "\n")
                    for i in range(len(self.results["Goal"])):
                        f.write(f"{i}," + ",".join(f"{self.results[k][i]:.2f}" for k in RESULT_TYPES) + "\n")
        return self.results["Goal"], self.results["Out"], self.results["Timeout"]
