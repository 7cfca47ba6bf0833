from cutie_b200.model.object_summarizer import ObjectSummarizer  # noqa: F401
