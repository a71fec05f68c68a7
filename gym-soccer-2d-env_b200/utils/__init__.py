"""The helper modules the reference's training scripts import (`utils.logger_utils`, `utils.info_collector_callback`),
so that those scripts run against this package unchanged."""
