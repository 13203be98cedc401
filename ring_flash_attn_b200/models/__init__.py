"""Model-side integration: Hugging Face adapter and a small Llama-style attention block."""
